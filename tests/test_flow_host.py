"""CPU execution of the velocity-side kernels of csrc/hdg_flow.cuh (BDM projection `common.py:91-108`, weak
divergence `hdg_imex.py:353-365` / `hdg_implicit.py:145`, pressure gradient `hdg_imex.py:333-340`, trace
reconstruction `hdg_imex.py:450-469`), compiled with g++ through tests/host_kernels (test infrastructure; the
engine has no CPU path), against the oracle.  The GPU twin is tests/test_engine_flow_gpu.py."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from oracle.hdg_oracle import HDGOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402
from test_poisson_host import HostMesh, dp, ip, rel  # noqa: E402

MESHES = [lambda: UnitSquareMesh(4, perturb=0.2), lambda: PeriodicSquareMesh(3, L=2 * np.pi), lambda: UnitDiskMesh(1)]
cd = ctypes.c_double


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return host_build.build("flow_host.cpp", str(tmp_path_factory.mktemp("host_kernels")))


def soa_Q(Q):
    return np.ascontiguousarray(Q.transpose(1, 2, 0).reshape(-1, Q.shape[0]))


def aos_Q(Qs, nq1):
    return Qs.reshape(2, nq1, -1).transpose(2, 0, 1)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("mesh_fn", MESHES)
def test_flow_kernels_on_the_host(lib, k, mesh_fn):
    mesh = mesh_fn()
    hm, o = HostMesh(mesh), HDGOracle(mesh, k)
    nc, nf, nq1 = mesh.nc, mesh.nf, o.nQ1
    rng = np.random.default_rng(k)
    Q = rng.standard_normal((nc, 2, nq1))
    p, lam = rng.standard_normal((nc, o.np_)), rng.standard_normal((nf, k + 1))
    Qs_, ps_, ls_ = soa_Q(Q), np.ascontiguousarray(p.T), np.ascontiguousarray(lam.T)
    nbr, nbr_e = np.zeros((3, nc), np.int32), np.zeros((3, nc), np.int32)
    assert lib.fh_build_nbr(nc, nf, ip(hm.cell_facet), ip(hm.facet_cell), ip(hm.facet_local), ip(nbr), ip(nbr_e)) == 0
    assert np.array_equal(nbr.T, o.nbr)

    # BDM projection
    fm, out = np.zeros((2 * (k + 2), nf)), np.zeros_like(Qs_)
    assert lib.fh_project_bdm(k, nc, nf, dp(hm.xy), ip(hm.cell_facet), ip(hm.facet_cell), dp(Qs_), dp(fm), dp(out)) == 0
    assert rel(aos_Q(out, nq1), o.project_bdm(Q)) < 1e-11

    # weak divergence, both right-hand-side families
    Rp = np.zeros((o.np_, nc))
    assert lib.fh_weak_div(k, nc, dp(hm.xy), ip(nbr), ip(nbr_e), dp(Qs_), cd(-2.5), 1, dp(Rp)) == 0
    assert rel(Rp.T, -2.5 * o.weak_divergence(Q)) < 1e-11
    assert lib.fh_weak_div(k, nc, dp(hm.xy), ip(nbr), ip(nbr_e), dp(Qs_), cd(0.5), 0, dp(Rp)) == 0
    assert rel(Rp.T, 0.5 * o.cell_divergence(Q)) < 1e-11

    # pressure gradient (Riesz form)
    Y = np.zeros_like(Qs_)
    assert lib.fh_pgrad(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), dp(ps_), dp(ls_), cd(0.0), cd(1.0),
                        dp(Y)) == 0
    assert rel(aos_Q(Y, nq1), o.pressure_gradient(p, lam) / o.detJ[:, None, None]) < 1e-11

    # trace reconstruction
    gK, lout = np.zeros((3 * (k + 1), nc)), np.zeros((k + 1, nf))
    assert lib.fh_reconstruct_trace(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.facet_cell), ip(hm.facet_local),
                                    cd(1.0), dp(Qs_), dp(ps_), dp(gK), dp(lout)) == 0
    assert rel(lout.T, o.reconstruct_trace(Q, p)) < 1e-11


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("mesh_fn", MESHES[:2])
def test_gamma_rows_and_reconstruction_rhs_on_the_host(lib, k, mesh_fn):
    """`k_gamma_cell` + `k_facet_sum` (the constraint rows of the monolithic operator, `hdg_imex.py:342-351`) against
    the oracle's assembled monolithic matrix, and `k_recon_rhs` (`hdg_imex.py:204-207`) against the oracle's
    quadrature restatement"""
    from oracle.timesteppers import IMEXOracle

    mesh = mesh_fn()
    ts = IMEXOracle(mesh, k, 0.1)
    hm, o = HostMesh(mesh), ts.o
    nc, nf, nq1 = mesh.nc, mesh.nf, o.nQ1
    rng = np.random.default_rng(10 + k)
    Q, B = rng.standard_normal((nc, 2, nq1)), rng.standard_normal((nc, 2, nq1))
    p, lam = rng.standard_normal((nc, o.np_)), rng.standard_normal((nf, k + 1))
    Qs_, Bs_, ps_, ls_ = soa_Q(Q), soa_Q(B), np.ascontiguousarray(p.T), np.ascontiguousarray(lam.T)

    Rp, gK, Rl = np.zeros((o.np_, nc)), np.zeros((3 * (k + 1), nc)), np.zeros((k + 1, nf))
    assert lib.fh_gamma(k, nc, nf, dp(hm.xy), ip(hm.cell_flip), ip(hm.cell_facet), ip(hm.facet_cell), ip(hm.facet_local),
                        cd(1.0), dp(Qs_), dp(ps_), dp(ls_), dp(Rp), dp(gK), dp(Rl)) == 0
    Kmat, (offp, offl, N) = o.assemble_monolithic()
    y = Kmat @ np.concatenate([Q.ravel(), p.ravel(), lam.ravel()])
    assert rel(Rp.T, y[offp:offl].reshape(nc, o.np_)) < 1e-11
    assert rel(Rl.T, y[offl:].reshape(nf, k + 1)) < 1e-11

    nbr, nbr_e = np.zeros((3, nc), np.int32), np.zeros((3, nc), np.int32)
    assert lib.fh_build_nbr(nc, nf, ip(hm.cell_facet), ip(hm.facet_cell), ip(hm.facet_local), ip(nbr), ip(nbr_e)) == 0
    Rp2, Rl2 = np.zeros((o.np_, nc)), np.zeros((k + 1, nf))
    assert lib.fh_recon_rhs(k, nc, nf, dp(hm.xy), ip(nbr), ip(nbr_e), ip(hm.cell_facet), ip(hm.cell_flip), dp(Qs_),
                            dp(Bs_), dp(Rp2), dp(Rl2)) == 0
    Rp_o, Rl_o = ts.reconstruction_rhs(Q, B)
    assert rel(Rp2.T, Rp_o) < 1e-11
    if np.abs(Rl_o).max() > 0:
        assert rel(Rl2.T, Rl_o) < 1e-11
    else:
        assert np.abs(Rl2).max() == 0.0
