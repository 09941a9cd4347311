"""CPU model of the tentative-velocity Krylov solve: BiCGStab iteration counts for preconditioner variants,
on the oracle's matrices (numpy/scipy; development tool kept under tests/ because it uses the oracle; not collected by pytest).

    python tests/experiments/tent_precond_model.py [nx=12] [k=2] [cfl=0.32]

System:  A x = b,  A = M - a f_impl(.;Q*)  (hdg_imex.py:233-235, hdg_implicit.py:103-125), a = cfl / nx,
Q* = BDM projection of the Taylor-Green velocity, split as  A = M - a F0 + a Pen  with the advection part F0
(alpha = 0) and the normal-jump penalty Pen = alpha N^T N.

  P1   (M + a Pen)^-1                         what the engine's facet-multiplier preconditioner (csrc/hdg_tent.cuh)
                                              applies when its Chebyshev sweeps are converged
  P8   (M + a Pen)^-1 M blockdiag(M - a F0)^-1   P1 composed with the inverse cell-diagonal advection blocks
                                              = the engine's experimental knob "tent_cellblock" (csrc/hdg_advblock.cuh)
  P3   (blockdiag(M - a F0) + a Pen)^-1       the combined operator (not cheaply invertible matrix-free; the bound)
  P12  (M + a Pen)^-1 M (M - a F0)^-1         P1 composed with the exact advection inverse

Measured here (nx, k = 2, cfl 0.32; iterations to rtol 1e-12, rough right-hand side / smooth warm-started):
  nx = 12:  P1 88 / 76,  P8 48 / 38,  P3 27 / 24,  P12 50 / 47
  nx = 24:  P1 96 / 76,  P8 52 / 42,  P3 28 / 24,  P12 60 / 54
and with six digits to gain from a warm start (the regime of bench.py): P1 31, P8 16, P3 9.
So the cell blocks halve the iteration count, independently of h; the product form cannot do better than that
(P12), only the combined operator P3 could.
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from oracle.hdg_oracle import HDGOracle  # noqa: E402
from oracle.timesteppers import TaylorGreenOracle  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfl = float(sys.argv[3]) if len(sys.argv) > 3 else 0.32
    adt = cfl / nx
    mesh = UnitSquareMesh(nx, perturb=0.1)
    t0 = time.time()
    o1, o0 = HDGOracle(mesh, k, alpha_penalty=1.0), HDGOracle(mesh, k, alpha_penalty=0.0)
    prob = TaylorGreenOracle("exponential", 0.5)
    Q0 = o1.interpolate_cell(lambda x, y: prob.Q_stationary(x, y), "Q")
    Qs = o1.project_bdm(Q0)
    F1, F0 = o1.f_impl_matrix(Qs), o0.f_impl_matrix(Qs)
    nc, nQ = mesh.nc, o1.nQ
    n = nc * nQ
    M = sp.diags(np.repeat(o1.detJ, nQ))
    Pen = (F0 - F1).tocsr()  # alpha N^T N, positive semidefinite
    A = (M - adt * F1).tocsc()
    print(f"nx={nx} k={k} n={n} a={adt:.4g} (matrices built in {time.time() - t0:.1f} s)", flush=True)

    B0 = (M - adt * F0).tobsr(blocksize=(nQ, nQ))
    D = np.zeros((nc, nQ, nQ))
    for i in range(nc):
        for jj in range(B0.indptr[i], B0.indptr[i + 1]):
            if B0.indices[jj] == i:
                D[i] = B0.data[jj]
    Dinv = np.linalg.inv(D)

    def cellblock(r):
        return np.einsum("nij,nj->ni", Dinv, r.reshape(nc, nQ)).ravel()

    Ppen = spla.splu((M + adt * Pen).tocsc())
    Bblk = sp.bsr_matrix((D, np.arange(nc), np.arange(nc + 1)), shape=(n, n)).tocsr()
    Pcomb = spla.splu((Bblk + adt * Pen).tocsc())
    B0lu = spla.splu(B0.tocsc())
    variants = [
        ("P1  penalty only (engine default)", Ppen.solve),
        ("P8  P1 * M * cellblock(M - a F0)^-1 (knob)", lambda r: Ppen.solve(M @ cellblock(r))),
        ("P3  (cellblock(M - a F0) + a Pen)^-1", Pcomb.solve),
        ("P12 P1 * M * (M - a F0)^-1", lambda r: Ppen.solve(M @ B0lu.solve(r))),
    ]
    rng = np.random.default_rng(0)
    b_rough = M @ Q0.ravel() + 1e-3 * (M @ rng.standard_normal(n))
    xs = Q0.ravel() * (1 - 0.5 * adt)
    b_smooth = A @ xs + adt * adt * (M @ np.sin(7 * np.arange(n) / n))
    cases = [("rough rhs, zero guess, rtol 1e-12", b_rough, None, 1e-12),
             ("smooth rhs, warm start, rtol 1e-12", b_smooth, Q0.ravel().copy(), 1e-12),
             ("smooth rhs, warm start, rtol 1e-6", b_smooth, Q0.ravel().copy(), 1e-6)]
    for label, b, x0, rtol in cases:
        print(label)
        for name, prec in variants:
            count = [0]
            x, info = spla.bicgstab(A, b, x0=x0, rtol=rtol, atol=0, maxiter=500,
                                    M=spla.LinearOperator((n, n), matvec=prec),
                                    callback=lambda xk: count.__setitem__(0, count[0] + 1))
            res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
            print(f"  {name:46s} iterations {count[0]:4d}  true relative residual {res:.1e}", flush=True)


if __name__ == "__main__":
    main()
