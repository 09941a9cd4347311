"""The C-ABI library builds for sm_100a, loads without a GPU and exports every declared symbol."""
import ctypes
import os
import re

from incompressibleeulerhdg_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(engine_lib):
    hdr = open(os.path.join(ROOT, "include", "hdg_b200.h")).read()
    declared = set(re.findall(r"\b(hdg_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"hdg_comm_init"} if "int hdg_comm_init" not in hdr else set()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(engine_lib, name), f"{name} declared in include/hdg_b200.h but not exported"
    # the Python binding table covers the same set
    assert declared == set(engine.SIGNATURES), declared ^ set(engine.SIGNATURES)


def test_version_and_degrees(engine_lib):
    assert b"sm_100a" in engine_lib.hdg_version()
    deg = engine_lib.hdg_supported_degrees()
    want = [int(k) for k in os.environ.get("HDG_DEV_DEGREES", "1,2,3,4").split(",")]
    assert all(deg & (1 << k) for k in want)


def test_create_rejects_bad_arguments(engine_lib):
    h = ctypes.c_void_p()
    rc = engine_lib.hdg_create(7, 1.0, 1, 1, None, None, None, None, None, 0, ctypes.byref(h))
    assert rc == engine.HDG_EINVAL
    assert b"degree" in engine_lib.hdg_last_error(None)


def test_no_cpu_fallback(engine_lib):
    """without a GPU the product path must fail loudly (no silent CPU route)"""
    import pytest

    from incompressibleeulerhdg_b200.mesh import UnitSquareMesh

    if engine_lib.hdg_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(engine.HDGError) as ei:
        engine.HDGEngine(UnitSquareMesh(2), 1)
    assert ei.value.code == engine.HDG_ENOGPU


def test_tuning_spec_parser():
    """HDG_TUNING, the environment switch for A/B runs of the unchanged tests and bench"""
    import pytest

    from incompressibleeulerhdg_b200.engine import parse_tuning

    assert parse_tuning("") == []
    assert parse_tuning("tent_cellblock") == [("tent_cellblock", 1)]
    assert parse_tuning(" tent_cellblock=1, tent_sweeps = 4 ,") == [("tent_cellblock", 1), ("tent_sweeps", 4)]
    with pytest.raises(ValueError):
        parse_tuning("tent_sweeps=four")
