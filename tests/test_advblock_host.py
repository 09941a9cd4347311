"""CPU check of the device code of the cell-block advection preconditioner (csrc/hdg_advblock.cuh, the
experimental ``tent_cellblock`` knob): the per-cell bodies are compiled with g++ (tests/host_kernels, CUDA
qualifiers defined away) and compared with the diagonal blocks of the oracle's f_impl matrix
(`hdg_imex.py:313-331` with alpha = 0).  This is test infrastructure: the engine itself has no CPU path."""
import ctypes
import os
import sys

import numpy as np
import pytest

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import TaylorGreenOracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_kernels"))
import build as host_build  # noqa: E402


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    lib = host_build.build("advblock_host.cpp", str(tmp_path_factory.mktemp("host_kernels")))
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    lib.advblock_host.restype = ctypes.c_int
    lib.advblock_host.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ip, dp, ctypes.c_double, dp,
                                  ctypes.POINTER(ctypes.c_float), dp, dp]
    return lib


def _ptr(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def _setup(k, nx, flux, noise=0.05):
    mesh = UnitSquareMesh(nx, perturb=0.15)
    o = HDGOracle(mesh, k, alpha_penalty=0.0, flux=flux)
    prob = TaylorGreenOracle("exponential", 0.5)
    rng = np.random.default_rng(5)
    Q = o.interpolate_cell(lambda x, y: prob.Q_stationary(x, y), "Q")
    Qs = o.project_bdm(Q + noise * rng.standard_normal(Q.shape))  # both signs of Q*.n on the facets
    xy = np.ascontiguousarray(np.asarray(mesh.cell_xy).transpose(1, 2, 0).reshape(6, mesh.nc))
    nbr = np.ascontiguousarray(np.asarray(o.nbr).T, dtype=np.int32)
    Qsoa = np.ascontiguousarray(Qs.transpose(1, 2, 0).reshape(2 * o.nQ1, mesh.nc))
    return mesh, o, Qs, xy, nbr, Qsoa


@pytest.mark.parametrize("flux", ["upwind", "centered"])
@pytest.mark.parametrize("k,nx", [(1, 4), (2, 3), (3, 2)])
def test_blocks_match_the_oracle_matrix(host_lib, k, nx, flux):
    mesh, o, Qs, xy, nbr, Qsoa = _setup(k, nx, flux)
    nc, n1 = mesh.nc, o.nQ1
    adt = 0.4 / nx
    blk = np.zeros((n1 * n1, nc))
    assert host_lib.advblock_host(k, int(flux == "upwind"), nc, _ptr(xy), _ptr(nbr, ctypes.c_int32), _ptr(Qsoa), adt,
                                  _ptr(blk), None, None, None) == 0
    F0 = o.f_impl_matrix(Qs).toarray().reshape(nc, 2, n1, nc, 2, n1)
    for cell in range(nc):
        ref = np.eye(n1) - adt * F0[cell, 0, :, cell, 0, :] / o.detJ[cell]
        assert np.abs(F0[cell, 0, :, cell, 1, :]).max() < 1e-13  # alpha = 0: the components decouple
        assert np.abs(F0[cell, 1, :, cell, 1, :] - F0[cell, 0, :, cell, 0, :]).max() < 1e-12
        got = blk[:, cell].reshape(n1, n1)
        assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()), (cell, np.abs(got - ref).max())


@pytest.mark.parametrize("k,nx", [(1, 4), (2, 3), (3, 2), (4, 2)])
def test_inverse_and_apply(host_lib, k, nx):
    mesh, o, Qs, xy, nbr, Qsoa = _setup(k, nx, "upwind", noise=0.0)  # the (nearly solenoidal) Q* of a real step
    nc, n1 = mesh.nc, o.nQ1
    adt = 0.32 / nx
    args = (k, 1, nc, _ptr(xy), _ptr(nbr, ctypes.c_int32), _ptr(Qsoa), adt)
    blk = np.zeros((n1 * n1, nc))
    assert host_lib.advblock_host(*args, _ptr(blk), None, None, None) == 0
    rng = np.random.default_rng(7)
    X = rng.standard_normal((2 * n1, nc))
    Y = np.zeros_like(X)
    inv, inv32 = np.zeros_like(blk), np.zeros(blk.shape, np.float32)
    assert host_lib.advblock_host(*args, _ptr(inv), _ptr(inv32, ctypes.c_float), _ptr(X), _ptr(Y)) == 0
    B = blk.T.reshape(nc, n1, n1)
    C = inv.T.reshape(nc, n1, n1)
    assert np.abs(np.einsum("nij,njk->nik", C, B) - np.eye(n1)).max() < 1e-11
    assert np.array_equal(inv32, inv.astype(np.float32))  # the copy the apply kernel reads (FP32 storage, FP64 arithmetic)
    C = inv32.astype(np.float64).T.reshape(nc, n1, n1)
    # positive definite symmetric part (I + a int_dK |s| phi phi - a/2 int_K div(Q*) phi phi with div Q* ~ 0): the
    # reason elimination without pivoting is safe
    assert np.linalg.eigvalsh(0.5 * (B + B.transpose(0, 2, 1))).min() > 0.0
    Xc = X.reshape(2, n1, nc)
    ref = np.einsum("nij,cjn->cin", C, Xc).reshape(2 * n1, nc)
    assert np.abs(Y - ref).max() < 1e-12 * np.abs(ref).max()
