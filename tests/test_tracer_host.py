"""CPU execution of the passive-tracer kernels of csrc/hdg_tracer.cuh (`common.py:110-129`,
`hdg_imex.py:415-448`), compiled with g++ through tests/host_kernels (test infrastructure; the engine has no CPU
path): the matrix-free Jacobi-PCG projection of the velocity onto [CG_{k+1}]^2 (launch sequence of
`run_project_cg`, csrc/hdg_engine.cu) and both advection kernels, against oracle/tracer.py.  The GPU twin is
tests/test_tracer_gpu.py."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200 import cgspace
from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.tracer import TracerOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402
from test_poisson_host import HostMesh, dp, ip, rel  # noqa: E402

cd = ctypes.c_double
TOL = 1e-10


def tg_velocity(x, y):
    return (-np.cos((x - 0.5) * np.pi) * np.sin((y - 0.5) * np.pi), np.sin((x - 0.5) * np.pi) * np.cos((y - 0.5) * np.pi))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return host_build.build("tracer_host.cpp", str(tmp_path_factory.mktemp("host_kernels")))


def soa_Q(Q):
    return np.ascontiguousarray(Q.transpose(1, 2, 0).reshape(-1, Q.shape[0]))


def aos_Q(Qs, nq1):
    return Qs.reshape(2, nq1, -1).transpose(2, 0, 1)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(4, perturb=0.15), lambda: UnitDiskMesh(1)])
def test_cg_projection_on_the_host(lib, k, mesh_fn):
    mesh = mesh_fn()
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    sp_ = cgspace.build_cg_space(mesh, k + 1)
    cellmap = np.ascontiguousarray(sp_.cellmap.T, dtype=np.int32)
    inc_ptr, inc_idx = np.ascontiguousarray(sp_.inc_ptr, np.int32), np.ascontiguousarray(sp_.inc_idx, np.int32)
    dinv = np.ascontiguousarray(1.0 / sp_.diag)
    Q = o.interpolate_cell(tg_velocity, "Q") + 0.1 * np.random.default_rng(2).standard_normal((mesh.nc, 2, o.nQ1))
    out, its = np.zeros((2 * o.nQ1, mesh.nc)), ctypes.c_int(0)
    rc = lib.trh_project_cg(k, mesh.nc, sp_.ndof, dp(hm.xy), ip(cellmap), ip(inc_ptr), ip(inc_idx), dp(dinv),
                            dp(soa_Q(Q)), dp(out), cd(1e-14), 300, ctypes.byref(its))
    assert rc == 0 and 0 < its.value < 200
    assert rel(aos_Q(out, o.nQ1), TracerOracle(o).project_cg(Q)) < TOL


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("mesh_fn", [lambda: UnitSquareMesh(4, perturb=0.15), lambda: PeriodicSquareMesh(3, L=2 * np.pi),
                                     lambda: UnitDiskMesh(1)])
def test_tracer_advection_on_the_host(lib, k, mesh_fn):
    mesh = mesh_fn()
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    nc = mesh.nc
    nbr = np.ascontiguousarray(np.asarray(o.nbr).T, dtype=np.int32)
    nbr_e = np.ascontiguousarray(np.asarray(o.nbr_e).T, dtype=np.int32)
    rng = np.random.default_rng(4)
    U = o.interpolate_cell(tg_velocity, "Q") + 0.1 * rng.standard_normal((nc, 2, o.nQ1))
    q, acc = rng.standard_normal((nc, o.np_)), rng.standard_normal((nc, o.np_))
    adv = TracerOracle(o).advection(q, U)
    Us, qs, accs = soa_Q(U), np.ascontiguousarray(q.T), np.ascontiguousarray(acc.T)
    tab_cell, tab_facet = cgspace.tracer_tables(k, o.nq_facet)
    for variant in ("compile-time tables", "runtime tables"):
        def run(c0, a, c1):
            out = np.zeros((o.np_, nc))
            if variant == "compile-time tables":
                rc = lib.trh_advect_t(k, nc, dp(hm.xy), ip(nbr), ip(nbr_e), dp(Us), dp(qs), cd(c0), dp(a), cd(c1), dp(out))
            else:
                rc = lib.trh_advect(k, nc, dp(hm.xy), ip(nbr), ip(nbr_e), tab_cell.shape[0], dp(tab_cell),
                                    tab_facet.shape[1], dp(tab_facet), dp(Us), dp(qs), cd(c0), dp(a), cd(c1), dp(out))
            assert rc == 0
            return out.T
        assert rel(run(0.0, None, 1.0), adv) < TOL, variant
        assert rel(run(0.5, accs, -0.25), 0.5 * acc - 0.25 * adv) < TOL, variant
