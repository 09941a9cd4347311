"""CPU model: Krylov methods x preconditioner variants for the tentative-velocity system at growing CFL
(development tool; uses the oracle, not collected by pytest).

    python tests/experiments/tent_cfl_model.py [nx=12] [k=2] [cfl list...]
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


def bicgstab_l(A, b, prec, l=2, rtol=1e-12, maxit=400):
    """BiCGStab(l) (Sleijpen-Fokkema), right preconditioned; returns (x, matvecs, relres)"""
    n = b.size
    op = lambda v: A @ prec(v)
    x = np.zeros(n)
    r = np.zeros((l + 1, n)); u = np.zeros((l + 1, n))
    r[0] = b
    rt = b.copy()
    bn = np.linalg.norm(b)
    rho0, alpha, omega = 1.0, 0.0, 1.0
    mv = 0
    while mv < maxit:
        rho0 = -omega * rho0
        for j in range(l):
            rho1 = r[j] @ rt
            beta = alpha * rho1 / rho0
            rho0 = rho1
            for i in range(j + 1):
                u[i] = r[i] - beta * u[i]
            u[j + 1] = op(u[j]); mv += 1
            gamma = u[j + 1] @ rt
            alpha = rho0 / gamma
            for i in range(j + 1):
                r[i] = r[i] - alpha * u[i + 1]
            r[j + 1] = op(r[j]); mv += 1
            x = x + alpha * u[0]
        # MR part
        R = r[1:l + 1]
        G = R @ R.T
        g = R @ r[0]
        gam = np.linalg.solve(G, g)
        omega = gam[-1]
        for j in range(1, l + 1):
            u[0] -= gam[j - 1] * u[j]
            x += gam[j - 1] * r[j - 1]
            r[0] -= gam[j - 1] * r[j]
        if np.linalg.norm(r[0]) <= rtol * bn:
            break
    xx = prec(x)
    return xx, mv, np.linalg.norm(b - A @ xx) / bn


def gmres_r(A, b, prec, m=30, rtol=1e-12, maxit=600):
    n = b.size
    bn = np.linalg.norm(b)
    x = np.zeros(n)
    mv = 0
    while mv < maxit:
        r = b - A @ x
        beta = np.linalg.norm(r)
        if beta <= rtol * bn:
            break
        V = np.zeros((m + 1, n)); Z = np.zeros((m, n)); H = np.zeros((m + 1, m))
        V[0] = r / beta
        jj = 0
        for j in range(m):
            Z[j] = prec(V[j])
            w = A @ Z[j]; mv += 1
            for i in range(j + 1):
                H[i, j] = w @ V[i]; w -= H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V[j + 1] = w / H[j + 1, j]
            jj = j + 1
            e1 = np.zeros(jj + 1); e1[0] = beta
            y, res, *_ = np.linalg.lstsq(H[:jj + 1, :jj], e1, rcond=None)
            rn = np.linalg.norm(e1 - H[:jj + 1, :jj] @ y)
            if rn <= rtol * bn or mv >= maxit:
                break
        x = x + Z[:jj].T @ y
    return x, mv, np.linalg.norm(b - A @ x) / bn


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfls = [float(a) for a in sys.argv[3:]] or [0.32, 1.0, 3.2, 10.0]
    mesh = UnitSquareMesh(nx, perturb=0.1)
    o1, o0 = HDGOracle(mesh, k, alpha_penalty=1.0), HDGOracle(mesh, k, alpha_penalty=0.0)
    prob = TaylorGreenOracle("exponential", 0.5)
    Q0 = o1.interpolate_cell(lambda x, y: prob.Q_stationary(x, y), "Q")
    Qs = o1.project_bdm(Q0)
    F1, F0 = o1.f_impl_matrix(Qs), o0.f_impl_matrix(Qs)
    nc, nQ = mesh.nc, o1.nQ
    n = nc * nQ
    M = sp.diags(np.repeat(o1.detJ, nQ))
    Pen = (F0 - F1).tocsr()
    rng = np.random.default_rng(0)
    for cfl in cfls:
        adt = cfl / nx
        A = (M - adt * F1).tocsc()
        B0 = (M - adt * F0).tobsr(blocksize=(nQ, nQ))
        D = np.zeros((nc, nQ, nQ))
        for i in range(nc):
            for jj in range(B0.indptr[i], B0.indptr[i + 1]):
                if B0.indices[jj] == i:
                    D[i] = B0.data[jj]
        Dinv = np.linalg.inv(D)
        cellblock = lambda r: np.einsum("nij,nj->ni", Dinv, r.reshape(nc, nQ)).ravel()
        Ppen = spla.splu((M + adt * Pen).tocsc())
        Bblk = sp.bsr_matrix((D, np.arange(nc), np.arange(nc + 1)), shape=(n, n)).tocsr()
        Pcomb = spla.splu((Bblk + adt * Pen).tocsc())
        variants = [("P1 ", Ppen.solve), ("P8 ", lambda r: Ppen.solve(M @ cellblock(r))), ("P3 ", Pcomb.solve)]
        b = M @ Q0.ravel() + 1e-3 * (M @ rng.standard_normal(n))
        print(f"nx={nx} k={k} cfl={cfl} n={n}", flush=True)
        for name, prec in variants:
            out = []
            cnt = [0]
            x, info = spla.bicgstab(A, b, rtol=1e-12, atol=0, maxiter=600, M=spla.LinearOperator((n, n), matvec=prec),
                                    callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
            out.append(f"bicgstab mv={2 * cnt[0]} res={np.linalg.norm(b - A @ x) / np.linalg.norm(b):.0e}")
            x, mv, res = bicgstab_l(A, b, prec, 2, maxit=1200)
            out.append(f"bicgstab(2) mv={mv} res={res:.0e}")
            for m in (20, 50):
                x, mv, res = gmres_r(A, b, prec, m, maxit=1200)
                out.append(f"gmres({m}) mv={mv} res={res:.0e}")
            print("  ", name, " | ".join(out), flush=True)


if __name__ == "__main__":
    main()
