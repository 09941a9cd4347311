import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine_lib():
    """build (if needed) and load the C-ABI library"""
    from incompressibleeulerhdg_b200 import build, engine

    build.build()
    return engine.load_library()


def require_degree(k):
    """skip when a development build (HDG_DEV_DEGREES) left this degree out"""
    from incompressibleeulerhdg_b200 import engine

    if not (engine.load_library().hdg_supported_degrees() >> k) & 1:
        pytest.skip(f"degree {k} not compiled into this build")
