"""Host-side *setup* of the geometric-trace multigrid hierarchy (numpy/scipy, setup time only).

The reference preconditions the trace system with ``firedrake.GTMGPC`` (`hdg_imex.py:138-169`): a
two-level method whose coarse space is conforming P1 on the same mesh (`get_coarse_space` :97-99,
`get_coarse_operator` :101-106 = -int grad.grad, constant null space :108-110), with the P1 problem
itself solved by one GAMG V-cycle (:153-167) and Chebyshev(2)/facet-block-Jacobi smoothing on the
trace level (:143-152).  This module builds the same ingredients as CSR matrices; every *apply*
(V-cycle, smoothers, transfers) then runs on the GPU inside the engine (``hdg_mg_setup``).

* transfer  T: P1 -> DGT_k is the exact trace of the P1 function (linear on each facet, so only
  Legendre modes 0 and 1 are populated).  The reference's ``interpolation_matrix`` (:491-503)
  carries a factor 1/2 on interior facets (``0.5*avg(u_coarse)``); that quirk only scales the
  coarse correction and is not reproduced.
* coarse operator: the P1 stiffness matrix.  It *equals* the Galerkin product T^T (-S) T of the
  condensed HDG operator to round-off (checked in tests), so rediscretisation and Galerkin agree.
* P1 hierarchy (stand-in for GAMG): nested geometric coarsening where the mesh generator knows
  one (structured squares, refined disk) with Galerkin coarse operators; the coarsest level is
  solved with a dense pseudo-inverse (the constant null space is projected out).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = ["p1_stiffness", "trace_transfer", "structured_prolongations", "build_hierarchy", "Hierarchy"]


def p1_stiffness(mesh) -> sp.csr_matrix:
    """int grad(phi_i).grad(phi_j) dx on the topological vertices of `mesh`"""
    x = mesh.cell_xy
    v = mesh.cell_vert.astype(np.int64)
    e = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)  # edge opposite vertex i
    area = mesh.cell_area()
    K = np.einsum("nid,njd->nij", e, e) / (4 * area[:, None, None])
    rows = np.repeat(v[:, :, None], 3, axis=2).ravel()
    cols = np.repeat(v[:, None, :], 3, axis=1).ravel()
    A = sp.coo_matrix((K.ravel(), (rows, cols)), shape=(mesh.nv, mesh.nv)).tocsr()
    A.sum_duplicates()
    return A


def trace_transfer(mesh, k: int) -> sp.csr_matrix:
    """T [(k+1) nf, nv] in the engine's SoA trace numbering (mode * nf + facet): the trace of a P1
    function u on facet f = (v0 -> v1) is (u0+u1)/2 * l_0 + (u1-u0)/(2 sqrt 3) * l_1"""
    nf = mesh.nf
    fv = mesh.facet_vert.astype(np.int64)
    f = np.arange(nf, dtype=np.int64)
    c1 = 1.0 / (2.0 * np.sqrt(3.0))
    rows = np.concatenate([f, f, nf + f, nf + f])
    cols = np.concatenate([fv[:, 0], fv[:, 1], fv[:, 0], fv[:, 1]])
    vals = np.concatenate([np.full(nf, 0.5), np.full(nf, 0.5), np.full(nf, -c1), np.full(nf, c1)])
    T = sp.coo_matrix((vals, (rows, cols)), shape=((k + 1) * nf, mesh.nv)).tocsr()
    return T


def structured_prolongations(nx: int, ny: int, diagonal: str, periodic: bool, min_n: int = 2):
    """linear-interpolation prolongations of the nested P1 spaces on an nx x ny grid of squares
    split along `diagonal`, from finest to coarsest, while both directions stay even"""
    out = []
    while nx % 2 == 0 and ny % 2 == 0 and min(nx, ny) // 2 >= min_n and (not periodic or min(nx, ny) // 2 >= 3):
        cx, cy = nx // 2, ny // 2
        if periodic:
            fid = lambda i, j: (j % ny) * nx + (i % nx)
            cid = lambda i, j: (j % cy) * cx + (i % cx)
            ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
            nfine, ncoarse = nx * ny, cx * cy
        else:
            fid = lambda i, j: j * (nx + 1) + i
            cid = lambda i, j: j * (cx + 1) + i
            ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
            nfine, ncoarse = (nx + 1) * (ny + 1), (cx + 1) * (cy + 1)
        ii, jj = ii.ravel(), jj.ravel()
        rows, cols, vals = [], [], []
        io, jo = ii % 2 == 1, jj % 2 == 1

        def add(mask, di0, dj0, di1, dj1):
            i, j = ii[mask], jj[mask]
            r = fid(i, j)
            for (di, dj) in ((di0, dj0), (di1, dj1)):
                rows.append(r)
                cols.append(cid((i + di) // 2, (j + dj) // 2))
                vals.append(np.full(r.size, 0.5))

        m = ~io & ~jo
        rows.append(fid(ii[m], jj[m]))
        cols.append(cid(ii[m] // 2, jj[m] // 2))
        vals.append(np.ones(m.sum()))
        add(io & ~jo, -1, 0, 1, 0)
        add(~io & jo, 0, -1, 0, 1)
        if diagonal == "left":  # coarse diagonal joins (I, J+1) and (I+1, J)
            add(io & jo, -1, 1, 1, -1)
        else:
            add(io & jo, -1, -1, 1, 1)
        P = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nfine, ncoarse)).tocsr()
        out.append(P)
        nx, ny = cx, cy
    return out


class Hierarchy:
    """A[l] (l = 0 finest P1 level), P[l]: level l+1 -> l, T: P1 level 0 -> trace, dense pinv of A[-1]"""

    def __init__(self, T, A, P, pinv, lmax):
        self.T, self.A, self.P, self.pinv, self.lmax = T, A, P, pinv, lmax

    @property
    def nlevels(self):
        return len(self.A)


def _lmax_jacobi(A, iters=100, seed=0):
    """power-iteration estimate of lambda_max(D^-1 A)"""
    d = A.diagonal()
    v = np.random.default_rng(seed).standard_normal(A.shape[0])
    lam = 1.0
    for _ in range(iters):
        v = (A @ v) / d
        lam = np.linalg.norm(v)
        v /= lam
    return float(lam)


def _pinv_semidefinite(A: np.ndarray, rtol: float = 1e-9) -> np.ndarray:
    """pseudo-inverse of the symmetric positive semi-definite coarsest operator with an explicit,
    generous cut-off.  The constant null vector shows up as an eigenvalue of relative size ~1e-16, which
    sits right at numpy's default pinv cut-off (1e-15): when it survived, the coarse solve amplified
    round-off along the constants by ~1e16, the preconditioned CG directions acquired huge null-space
    components and <p, P p> lost all accuracy (erratic, occasionally stalling trace solves;
    profiles/debug_cg_trace_r1o.log)."""
    A = 0.5 * (A + A.T)
    w, V = np.linalg.eigh(A)
    keep = w > rtol * w.max()
    assert keep.sum() >= A.shape[0] - 1 - 2, "coarsest operator has an unexpectedly large null space"
    return (V[:, keep] / w[keep]) @ V[:, keep].T


def build_hierarchy(mesh, k: int, max_coarsest: int = 1200) -> Hierarchy:
    T = trace_transfer(mesh, k)
    A0 = p1_stiffness(mesh)
    Ps = mesh.meta.get("p1_prolongations")
    if Ps is None:
        meta = mesh.meta
        if "nx" in meta:
            Ps = structured_prolongations(meta["nx"], meta["ny"], meta["diagonal"], meta["periodic"])
        else:
            Ps = []
    A, P = [A0], []
    for Pl in Ps:
        if A[-1].shape[0] <= 32:
            break
        assert Pl.shape[0] == A[-1].shape[0]
        Ac = (Pl.T @ A[-1] @ Pl).tocsr()
        Ac.sum_duplicates()
        A.append(Ac)
        P.append(Pl.tocsr())
    n_last = A[-1].shape[0]
    if n_last > max_coarsest:
        raise ValueError(
            f"coarsest P1 level has {n_last} unknowns (> {max_coarsest}): this mesh carries no nested hierarchy; "
            "algebraic coarsening is not implemented yet")
    pinv = _pinv_semidefinite(A[-1].toarray())
    lmax = [_lmax_jacobi(a) for a in A]
    return Hierarchy(T, A, P, pinv, lmax)
