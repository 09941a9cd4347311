"""CPU model: block-Jacobi / two-colour block Gauss-Seidel / lexicographic block Gauss-Seidel sweeps on the full
tentative-velocity operator as GMRES preconditioners at CFL 10-40 (development tool; uses the oracle, not collected by
pytest; result in profiles/r2/tent_gs_model_nx16.log).

    python tests/experiments/tent_gs_model.py nx k cfl...
"""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, __file__.rsplit("/tests/", 1)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import TaylorGreenOracle
from tent_cfl_model import gmres_r, bicgstab_l

nx = int(sys.argv[1]); k = int(sys.argv[2]); cfls = [float(a) for a in sys.argv[3:]]
mesh = UnitSquareMesh(nx, perturb=0.1)
o1 = HDGOracle(mesh, k, alpha_penalty=1.0)
prob = TaylorGreenOracle("exponential", 0.5)
Q0 = o1.interpolate_cell(lambda x, y: prob.Q_stationary(x, y), "Q")
Qs = o1.project_bdm(Q0)
F1 = o1.f_impl_matrix(Qs)
nc, nQ = mesh.nc, o1.nQ
n = nc * nQ
M = sp.diags(np.repeat(o1.detJ, nQ))
rng = np.random.default_rng(0)
for cfl in cfls:
    adt = cfl / nx
    A = (M - adt * F1).tocsr()
    Ab = A.tobsr(blocksize=(nQ, nQ))
    # cell graph
    G = sp.csr_matrix((np.ones(len(Ab.indices)), Ab.indices, Ab.indptr), shape=(nc, nc))
    deg = np.diff(G.indptr)
    print(f"nx={nx} k={k} cfl={cfl} n={n} maxdeg={deg.max()}", flush=True)
    # greedy colouring
    colour = -np.ones(nc, int)
    for i in range(nc):
        used = set(colour[G.indices[G.indptr[i]:G.indptr[i+1]]])
        c = 0
        while c in used: c += 1
        colour[i] = c
    ncol = colour.max() + 1
    print("   colours", ncol, np.bincount(colour))
    D = np.zeros((nc, nQ, nQ))
    for i in range(nc):
        for jj in range(Ab.indptr[i], Ab.indptr[i+1]):
            if Ab.indices[jj] == i: D[i] = Ab.data[jj]
    Dinv = np.linalg.inv(D)
    bj = lambda r: np.einsum("nij,nj->ni", Dinv, r.reshape(nc, nQ)).ravel()
    def mcgs(nsweep, sym=False):
        masks = [np.repeat(colour == c, nQ) for c in range(ncol)]
        def f(r):
            x = np.zeros(n)
            for s in range(nsweep):
                order = list(range(ncol))
                if sym and s % 2 == 1: order = order[::-1]
                for c in order:
                    res = r - A @ x
                    x[masks[c]] += bj(res)[masks[c]]
            return x
        return f
    def jac(nsweep, w=1.0):
        def f(r):
            x = np.zeros(n)
            for s in range(nsweep):
                x += w * bj(r - A @ x)
            return x
        return f
    # lexicographic block GS (sequential reference) via sparse triangular solve
    Lb = sp.tril(Ab.tocsr(), 0).tocsr()  # point-lower incl. diag -- approx; use block lower:
    perm = np.arange(nc)
    def blocktri(order):
        rank = np.empty(nc, int); rank[order] = np.arange(nc)
        rows, cols = Ab.tocoo().row, Ab.tocoo().col
        Ac = A.tocoo()
        keep = rank[Ac.row // nQ] >= rank[Ac.col // nQ]
        L = sp.csc_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=(n, n))
        return spla.splu(L, permc_spec="NATURAL") if False else spla.splu(L)
    Lnat = blocktri(np.arange(nc))
    def gs_nat(nsweep):
        def f(r):
            x = np.zeros(n)
            for s in range(nsweep):
                x += Lnat.solve(r - A @ x)
            return x
        return f
    variants = [("BJ1", jac(1), 1), ("BJ4", jac(4), 4), ("BJ8 w.8", jac(8, 0.8), 8),
                ("MCGS1", mcgs(1), 1), ("MCGS2", mcgs(2), 2), ("MCGS4", mcgs(4), 4), ("MCGS4s", mcgs(4, True), 4), ("MCGS8", mcgs(8), 8),
                ("GSnat1", gs_nat(1), 1), ("GSnat4", gs_nat(4), 4),
                ]
    b = M @ Q0.ravel() + 1e-3 * (M @ rng.standard_normal(n))
    for name, prec, cost in variants:
        t = time.time()
        x, mv, res = gmres_r(A, b, prec, 30, maxit=600)
        x2, mv2, res2 = bicgstab_l(A, b, prec, 1, maxit=600) if False else (None, 0, 0)
        print(f"   {name:12s} gmres(30) mv={mv} res={res:.0e} (operator applications ~{mv*(cost+1)})  {time.time()-t:.1f}s", flush=True)
