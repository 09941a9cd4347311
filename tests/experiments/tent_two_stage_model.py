"""CPU model: multiplicative two-stage preconditioners for the tentative-velocity system at CFL 10-40 -- an exact
transport solve T = (M - a F0)^-1 (no penalty) combined with the exact cell-block + penalty operator P3 or the penalty
operator P1 (development tool; uses the oracle, not collected by pytest; result in profiles/r2/tent_two_stage_model_nx16.log).

    python tests/experiments/tent_two_stage_model.py nx k cfl...
"""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, __file__.rsplit("/tests/", 1)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh
from oracle.hdg_oracle import HDGOracle
from oracle.timesteppers import TaylorGreenOracle
from tent_cfl_model import gmres_r

nx = int(sys.argv[1]); k = int(sys.argv[2]); cfls = [float(a) for a in sys.argv[3:]]
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
def blockdiag(Mat):
    Bb = Mat.tobsr(blocksize=(nQ, nQ))
    D = np.zeros((nc, nQ, nQ))
    for i in range(nc):
        for jj in range(Bb.indptr[i], Bb.indptr[i + 1]):
            if Bb.indices[jj] == i:
                D[i] = Bb.data[jj]
    return D
for cfl in cfls:
    adt = cfl / nx
    A = (M - adt * F1).tocsc()
    T0 = (M - adt * F0).tocsc()           # transport only (no penalty)
    D = blockdiag(T0)
    Bblk = sp.bsr_matrix((D, np.arange(nc), np.arange(nc + 1)), shape=(n, n)).tocsr()
    P3 = spla.splu((Bblk + adt * Pen).tocsc())
    P1 = spla.splu((M + adt * Pen).tocsc())
    Tlu = spla.splu(T0)
    def two(stage1, stage2):
        def f(r):
            x = stage1(r)
            return x + stage2(r - A @ x)
        return f
    def three(s1, s2):
        def f(r):
            x = s1(r); x = x + s2(r - A @ x); return x + s1(r - A @ x)
        return f
    b = M @ Q0.ravel() + 1e-3 * (M @ rng.standard_normal(n))
    print(f"nx={nx} k={k} cfl={cfl} n={n}", flush=True)
    for name, prec in [("P3", P3.solve), ("T", Tlu.solve), ("T then P3", two(Tlu.solve, P3.solve)), ("P3 then T", two(P3.solve, Tlu.solve)),
                       ("T then P1", two(Tlu.solve, P1.solve)), ("P1 then T", two(P1.solve, Tlu.solve)), ("T P3 T", three(Tlu.solve, P3.solve)), ("P3 T P3", three(P3.solve, Tlu.solve))]:
        t = time.time()
        x, mv, res = gmres_r(A, b, prec, 30, maxit=400)
        print(f"   {name:12s} gmres(30) mv={mv} res={res:.0e}  {time.time()-t:.1f}s", flush=True)
