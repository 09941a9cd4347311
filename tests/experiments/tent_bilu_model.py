"""CPU model: block ILU(0) on the cell-block matrix of the tentative-velocity system (the preconditioner class the
reference itself uses, GMRES + ILU `hdg_imex.py:224-228`) as a right preconditioner of GMRES(30) at CFL 10-40, for several
cell orderings (development tool; uses the oracle, not collected by pytest; result in profiles/r2/tent_bilu_model_nx16.log).

    python tests/experiments/tent_bilu_model.py nx k cfl...
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from oracle.hdg_oracle import HDGOracle  # noqa: E402
from oracle.timesteppers import TaylorGreenOracle  # noqa: E402
from tent_cfl_model import gmres_r  # noqa: E402


class BlockILU0:
    """ILU(0) on nb x nb blocks, rows eliminated in the given order; apply = forward + backward block substitution"""

    def __init__(self, A, nc, nb, order):
        self.nc, self.nb, self.order = nc, nb, order
        rank = np.empty(nc, int); rank[order] = np.arange(nc)
        Ab = A.tobsr(blocksize=(nb, nb))
        self.rows = [dict() for _ in range(nc)]  # in permuted numbering: rows[i][j] = block
        for i in range(nc):
            for jj in range(Ab.indptr[i], Ab.indptr[i + 1]):
                self.rows[rank[i]][rank[Ab.indices[jj]]] = Ab.data[jj].copy()
        self.dinv = [None] * nc
        for i in range(nc):
            row = self.rows[i]
            for kk in sorted(j for j in row if j < i):
                row[kk] = row[kk] @ self.dinv[kk]  # L_ik
                for j, ukj in self.rows[kk].items():
                    if j > kk and j in row:  # no fill outside the pattern
                        row[j] = row[j] - row[kk] @ ukj
            self.dinv[i] = np.linalg.inv(row[i])
        self.levels = self._levels()

    def _levels(self):
        lev = np.zeros(self.nc, int)
        for i in range(self.nc):
            lower = [j for j in self.rows[i] if j < i]
            lev[i] = 1 + max((lev[j] for j in lower), default=-1)
        return int(lev.max()) + 1

    def solve(self, r):
        nc, nb = self.nc, self.nb
        y = r.reshape(nc, nb)[self.order].copy()
        for i in range(nc):
            for j, blk in self.rows[i].items():
                if j < i:
                    y[i] -= blk @ y[j]
        for i in range(nc - 1, -1, -1):
            for j, blk in self.rows[i].items():
                if j > i:
                    y[i] -= blk @ y[j]
            y[i] = self.dinv[i] @ y[i]
        out = np.empty_like(y)
        out[self.order] = y
        return out.ravel()


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfls = [float(a) for a in sys.argv[3:]] or [10.0, 40.0]
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
    cen = np.asarray(mesh.cell_xy).mean(axis=1)  # [nc, 2] centroids
    for cfl in cfls:
        adt = cfl / nx
        A = (M - adt * F1).tocsr()
        G = A.tobsr(blocksize=(nQ, nQ))
        graph = sp.csr_matrix((np.ones(len(G.indices)), G.indices, G.indptr), shape=(nc, nc))
        orders = {
            "natural": np.arange(nc),
            "rcm": np.asarray(csgraph.reverse_cuthill_mckee(graph, symmetric_mode=False)),
            # the Taylor-Green vortex turns around the domain centre: order by angle (approximate downwind order)
            "angle": np.argsort(np.arctan2(cen[:, 1] - 0.5, cen[:, 0] - 0.5)),
            "two-colour": np.argsort(np.asarray([sum(1 for _ in ()) for _ in range(nc)]), kind="stable"),
        }
        # two-colour ordering of the bipartite cell graph (what a GPU would like best: 2 levels)
        colour = -np.ones(nc, int)
        for i in range(nc):
            used = set(colour[graph.indices[graph.indptr[i]:graph.indptr[i + 1]]])
            c = 0
            while c in used:
                c += 1
            colour[i] = c
        orders["two-colour"] = np.argsort(colour, kind="stable")
        b = M @ Q0.ravel() + 1e-3 * (M @ rng.standard_normal(n))
        print(f"nx={nx} k={k} cfl={cfl} n={n}", flush=True)
        for name, order in orders.items():
            t = time.time()
            P = BlockILU0(A, nc, nQ, order)
            x, mv, res = gmres_r(A, b, P.solve, 30, maxit=300)
            print(f"   block ILU(0), {name:10s} ordering ({P.levels:4d} dependency levels): gmres(30) mv={mv} res={res:.0e}"
                  f"  {time.time() - t:.0f}s", flush=True)


if __name__ == "__main__":
    main()
