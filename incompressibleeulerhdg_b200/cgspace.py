"""Continuous Lagrange space CG_d on a triangle mesh, and the quadrature tables of the tracer kernels.

The reference projects the advecting velocity onto ``VectorFunctionSpace(mesh, "CG", k+1)`` before it
forms the passive-tracer flux (`timesteppers/common.py:110-129`).  This module builds the *topology*
of that space -- the cell -> global dof map and its transpose -- which is all the engine needs: the
mass matrix is applied matrix-free (``csrc/hdg_tracer.cuh``).  Set-up is host-side numpy, once per run
(the analogue of Firedrake building a ``cell_node_map``).

Numbering (d = degree, nodes = ``refelem.lagrange_nodes_cell(d)``, barycentric indices (l, i, j) with
l + i + j = d for the vertices ((0,0), (1,0), (0,1))):

* vertex nodes      -> topological vertex id                                  [0, nv)
* facet-interior    -> nv + f (d-1) + (t-1), t = 1..d-1 counted from the facet's first vertex *in the
  nodes                global facet direction* (``cell_flip`` reverses the cell-local count)
* cell-interior     -> nv + nf (d-1) + cell * nint + running index
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import refelem as R

__all__ = ["CGSpace", "build_cg_space", "tracer_tables"]


@dataclass
class CGSpace:
    degree: int
    ndof: int
    cellmap: np.ndarray  # [nc, nloc] int32
    inc_ptr: np.ndarray  # [ndof + 1] int32
    inc_idx: np.ndarray  # [nc * nloc] int32, entries j * nc + cell sorted by (dof, cell, j)
    W: np.ndarray  # [nloc, nloc]  modal <- nodal
    diag: np.ndarray  # [ndof] diagonal of the mass matrix

    @property
    def nloc(self) -> int:
        return self.cellmap.shape[1]


def _local_entities(d: int):
    """classify the Lagrange nodes of P_d: list of ('v', vertex) | ('f', facet, t) | ('c', index)"""
    ents, nint = [], 0
    for j in range(d + 1):
        for i in range(d + 1 - j):
            bary = (d - i - j, i, j)
            zeros = [v for v in range(3) if bary[v] == 0]
            if len(zeros) == 2:
                ents.append(("v", int(np.argmax(bary))))
            elif len(zeros) == 1:
                e = zeros[0]  # facet e is opposite vertex e and runs (e+1)%3 -> (e+2)%3
                ents.append(("f", e, bary[(e + 2) % 3]))  # distance from the first vertex in units of 1/d
            else:
                ents.append(("c", nint))
                nint += 1
    return ents, nint


def build_cg_space(mesh, degree: int, perm: np.ndarray | None = None) -> CGSpace:
    """`perm` (optional) renumbers the dofs: dof l of the numbering above becomes perm[l] (used on a
    partitioned mesh to put the owned dofs first, ``partition.cg_plan``)"""
    d = int(degree)
    assert d >= 1
    nc, nf, nv = mesh.nc, mesh.nf, mesh.nv
    ents, nint = _local_entities(d)
    nloc = len(ents)
    cellmap = np.empty((nc, nloc), dtype=np.int64)
    cells = np.arange(nc)
    for j, ent in enumerate(ents):
        if ent[0] == "v":
            cellmap[:, j] = mesh.cell_vert[:, ent[1]]
        elif ent[0] == "f":
            _, e, t = ent
            tg = np.where(mesh.cell_flip[:, e] != 0, d - t, t)
            cellmap[:, j] = nv + mesh.cell_facet[:, e].astype(np.int64) * (d - 1) + (tg - 1)
        else:
            cellmap[:, j] = nv + nf * (d - 1) + cells * nint + ent[1]
    ndof = nv + nf * (d - 1) + nc * nint
    assert ndof < 2 ** 31 and nc * nloc < 2 ** 31
    if perm is not None:
        assert perm.shape == (ndof,)
        cellmap = np.asarray(perm, dtype=np.int64)[cellmap]
    flat = cellmap.ravel()  # index = cell * nloc + j
    order = np.argsort(flat, kind="stable")
    counts = np.bincount(flat, minlength=ndof)
    assert counts.min() >= 1, "unreferenced CG dof"
    inc_ptr = np.concatenate([[0], np.cumsum(counts)])
    cell_of, j_of = np.divmod(order, nloc)
    inc_idx = j_of * nc + cell_of
    W = R.nodal_to_modal_cell(d)
    detJ = 2.0 * mesh.cell_area()
    local_diag = np.einsum("ij,ij->j", W, W)  # (W^T W)_jj
    diag = np.bincount(flat, weights=(detJ[:, None] * local_diag[None, :]).ravel(), minlength=ndof)
    return CGSpace(degree=d, ndof=int(ndof), cellmap=cellmap.astype(np.int32), inc_ptr=inc_ptr.astype(np.int32),
                   inc_idx=inc_idx.astype(np.int32), W=np.ascontiguousarray(W), diag=diag)


def tracer_tables(k: int, nq_facet: int | None = None):
    """quadrature tables of ``k_tracer_adv`` for tracer degree k and velocity degree k+1:

    ``tab_cell [nq, 1 + 3 NP + 3 NQ1]``  weight, chi, d_xi chi, d_eta chi, psi, d_xi psi, d_eta psi; the
    weights are those of the reference triangle (sum w = 1/2): the Dubiner basis is orthonormal on T^
    and the physical mass matrix is detJ I, so  M^-1 int_K f chi dx = sum_q w_q f chi  without detJ.
    ``tab_facet [3, nqf, 1 + NP + NQ1]`` Gauss-Legendre weight on [0,1], chi, psi at local facet e.

    The volume integrand q div(chi u) has degree 3k; the facet integrand is only piecewise polynomial
    (|u.n|); nq_facet defaults to ceil((3k+4)/2), the engine-wide facet-rule default."""
    NP, NQ1 = R.ncell(k), R.ncell(k + 1)
    xq, wq = R.triangle_quadrature_gj(3 * k + 1)
    chi, dchi = R.dubiner(k, xq), R.dubiner_grad(k, xq)
    psi, dpsi = R.dubiner(k + 1, xq), R.dubiner_grad(k + 1, xq)
    tab_cell = np.concatenate([wq[:, None], chi.T, dchi[:, :, 0].T, dchi[:, :, 1].T, psi.T, dpsi[:, :, 0].T,
                               dpsi[:, :, 1].T], axis=1)
    assert tab_cell.shape == (len(wq), 1 + 3 * NP + 3 * NQ1)
    nqf = (3 * k + 5) // 2 if nq_facet is None else int(nq_facet)
    s, wf = R.gauss_legendre(nqf)
    assert np.allclose(s[::-1], 1.0 - s), "the facet rule must be symmetric"
    tab_facet = np.empty((3, nqf, 1 + NP + NQ1))
    for e in range(3):
        pts = R.facet_points(e, s)
        tab_facet[e, :, 0] = wf
        tab_facet[e, :, 1:1 + NP] = R.dubiner(k, pts).T
        tab_facet[e, :, 1 + NP:] = R.dubiner(k + 1, pts).T
    return np.ascontiguousarray(tab_cell), np.ascontiguousarray(tab_facet)
