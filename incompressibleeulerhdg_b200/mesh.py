"""Firedrake-free triangle meshes and the topology arrays the engine consumes.

Mirrors the three mesh constructors the reference driver uses (`driver.py:180-185`):
``UnitSquareMesh(nx, nx, quadrilateral=False)``, ``PeriodicSquareMesh(nx, nx, L)`` and (through
:meth:`Mesh.from_cells`) arbitrary unstructured affine triangle meshes such as ``UnitDiskMesh``.

Array schema (all C-contiguous; this is exactly what ``hdg_create`` in ``include/hdg_b200.h`` takes,
so a Firedrake adapter is a pure data-marshalling shim):

================  ===========  =====================================================================
``cell_xy``       [nc,3,2] f8  vertex coordinates *per cell* (so periodic meshes need no special case)
``cell_vert``     [nc,3]   i4  topological vertex ids, counter-clockwise
``cell_facet``    [nc,3]   i4  global facet id of local facet e (e is opposite local vertex e and
                               runs from local vertex (e+1)%3 to (e+2)%3)
``cell_flip``     [nc,3]   i4  1 if the cell traverses the facet against its global direction
``facet_cell``    [nf,2]   i4  adjacent cells, -1 in slot 1 on the boundary
``facet_local``   [nf,2]   i4  local facet index inside each adjacent cell (-1 if absent)
================  ===========  =====================================================================

The global direction of a facet is "from the smaller to the larger topological vertex id".
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

__all__ = ["Mesh", "UnitSquareMesh", "PeriodicSquareMesh", "UnitDiskMesh", "RandomAffineCells"]


@dataclass
class Mesh:
    cell_xy: np.ndarray
    cell_vert: np.ndarray
    cell_facet: np.ndarray
    cell_flip: np.ndarray
    facet_cell: np.ndarray
    facet_local: np.ndarray
    facet_vert: np.ndarray
    nv: int
    name: str = "mesh"
    meta: dict = field(default_factory=dict)

    @property
    def nc(self) -> int:
        return self.cell_xy.shape[0]

    @property
    def nf(self) -> int:
        return self.facet_cell.shape[0]

    @property
    def boundary_facets(self) -> np.ndarray:
        return np.nonzero(self.facet_cell[:, 1] < 0)[0]

    @property
    def volume(self) -> float:
        return float(self.cell_area().sum())

    def cell_area(self) -> np.ndarray:
        x = self.cell_xy
        d = (x[:, 1, 0] - x[:, 0, 0]) * (x[:, 2, 1] - x[:, 0, 1]) - (x[:, 2, 0] - x[:, 0, 0]) * (
            x[:, 1, 1] - x[:, 0, 1]
        )
        return 0.5 * d

    def facet_length(self) -> np.ndarray:
        """|F| for every facet, measured in the first adjacent cell (`common.py:36-57` stores 1/h_F)"""
        c = self.facet_cell[:, 0]
        e = self.facet_local[:, 0]
        a = self.cell_xy[c, (e + 1) % 3]
        b = self.cell_xy[c, (e + 2) % 3]
        return np.hypot(*(b - a).T)

    @staticmethod
    def from_cells(cell_vert: np.ndarray, vert_xy: np.ndarray | None = None, cell_xy: np.ndarray | None = None,
                   name: str = "mesh") -> "Mesh":
        """Build all topology arrays from a cell->vertex map (vectorised, O(nc log nc))."""
        cell_vert = np.ascontiguousarray(cell_vert, dtype=np.int64)
        nc = cell_vert.shape[0]
        if cell_xy is None:
            cell_xy = np.asarray(vert_xy, dtype=np.float64)[cell_vert]
        cell_xy = np.array(cell_xy, dtype=np.float64, copy=True)
        # make every cell counter-clockwise
        x = cell_xy
        det = (x[:, 1, 0] - x[:, 0, 0]) * (x[:, 2, 1] - x[:, 0, 1]) - (x[:, 2, 0] - x[:, 0, 0]) * (
            x[:, 1, 1] - x[:, 0, 1]
        )
        neg = det < 0
        if np.any(neg):
            cell_vert[neg] = cell_vert[neg][:, [0, 2, 1]]
            cell_xy[neg] = cell_xy[neg][:, [0, 2, 1]]
        nv = int(cell_vert.max()) + 1
        # local facet e: (e+1)%3 -> (e+2)%3
        va = cell_vert[:, [1, 2, 0]]
        vb = cell_vert[:, [2, 0, 1]]
        lo = np.minimum(va, vb)
        hi = np.maximum(va, vb)
        key = (lo * nv + hi).ravel()
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        # number facets by first appearance so that facet ids inherit the locality of the cell order
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        fid = rank[inv]
        nf = uniq.size
        cell_facet = fid.reshape(nc, 3)
        cell_flip = (va > vb).astype(np.int32)
        facet_vert = np.stack([uniq[order] // nv, uniq[order] % nv], axis=1)
        facet_cell = np.full((nf, 2), -1, dtype=np.int64)
        facet_local = np.full((nf, 2), -1, dtype=np.int64)
        flat_cell = np.repeat(np.arange(nc), 3)
        flat_loc = np.tile(np.arange(3), nc)
        # first occurrence -> slot 0, second -> slot 1
        srt = np.argsort(fid, kind="stable")
        f_sorted = fid[srt]
        is_first = np.ones(f_sorted.size, dtype=bool)
        is_first[1:] = f_sorted[1:] != f_sorted[:-1]
        slot = np.where(is_first, 0, 1)
        if np.any(np.bincount(fid, minlength=nf) > 2):
            raise ValueError("non-manifold mesh: a facet has more than two cells")
        facet_cell[f_sorted, slot] = flat_cell[srt]
        facet_local[f_sorted, slot] = flat_loc[srt]
        return Mesh(
            cell_xy=np.ascontiguousarray(cell_xy),
            cell_vert=np.ascontiguousarray(cell_vert, dtype=np.int32),
            cell_facet=np.ascontiguousarray(cell_facet, dtype=np.int32),
            cell_flip=np.ascontiguousarray(cell_flip, dtype=np.int32),
            facet_cell=np.ascontiguousarray(facet_cell, dtype=np.int32),
            facet_local=np.ascontiguousarray(facet_local, dtype=np.int32),
            facet_vert=np.ascontiguousarray(facet_vert, dtype=np.int32),
            nv=nv,
            name=name,
        )


def _square_cells(nx: int, ny: int, diagonal: str, periodic: bool):
    """cell->vertex map of an nx x ny grid of squares split into two triangles each"""
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    i = i.ravel()
    j = j.ravel()
    if periodic:
        vid = lambda a, b: (b % ny) * nx + (a % nx)
    else:
        vid = lambda a, b: b * (nx + 1) + a
    v00, v10, v01, v11 = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
    if diagonal == "left":
        # diagonal from top-left to bottom-right
        t0 = np.stack([v00, v10, v01], axis=1)
        t1 = np.stack([v10, v11, v01], axis=1)
        c0 = ((0, 0), (1, 0), (0, 1))
        c1 = ((1, 0), (1, 1), (0, 1))
    elif diagonal == "right":
        t0 = np.stack([v00, v10, v11], axis=1)
        t1 = np.stack([v00, v11, v01], axis=1)
        c0 = ((0, 0), (1, 0), (1, 1))
        c1 = ((0, 0), (1, 1), (0, 1))
    else:
        raise ValueError("diagonal must be 'left' or 'right'")
    cell_vert = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cell_vert[0::2] = t0
    cell_vert[1::2] = t1
    # integer corner coordinates per cell (unwrapped, so periodic cells keep their true shape)
    cij = np.empty((2 * nx * ny, 3, 2), dtype=np.float64)
    for loc, (a, b) in enumerate(c0):
        cij[0::2, loc, 0] = i + a
        cij[0::2, loc, 1] = j + b
    for loc, (a, b) in enumerate(c1):
        cij[1::2, loc, 0] = i + a
        cij[1::2, loc, 1] = j + b
    return cell_vert, cij


def UnitSquareMesh(nx: int, ny: int | None = None, quadrilateral: bool = False, diagonal: str = "left",
                   perturb: float = 0.0) -> Mesh:
    """2 nx ny triangles on [0,1]^2 (`driver.py:181`).

    ``perturb`` > 0 moves interior vertices by a smooth field of amplitude ``perturb * h`` so that
    no two cells are congruent; boundary vertices stay on the boundary.
    """
    assert not quadrilateral, "only triangles are supported (the reference passes quadrilateral=False)"
    ny = nx if ny is None else ny
    cell_vert, cij = _square_cells(nx, ny, diagonal, periodic=False)
    cell_xy = cij / np.array([nx, ny], dtype=np.float64)
    if perturb:
        x = cell_xy[..., 0].copy()
        y = cell_xy[..., 1].copy()
        bump = np.sin(np.pi * x) * np.sin(np.pi * y)
        cell_xy[..., 0] = x + perturb / nx * bump * np.sin(7.0 * x + 3.0 * y + 0.3)
        cell_xy[..., 1] = y + perturb / ny * bump * np.cos(5.0 * x - 4.0 * y + 0.1)
    m = Mesh.from_cells(cell_vert, cell_xy=cell_xy, name=f"UnitSquareMesh({nx},{ny})")
    m.meta.update(nx=nx, ny=ny, diagonal=diagonal, periodic=False, L=1.0, perturb=perturb)
    return m


def PeriodicSquareMesh(nx: int, ny: int | None = None, L: float = 1.0, quadrilateral: bool = False,
                       diagonal: str = "left") -> Mesh:
    """doubly periodic [0,L]^2 (`driver.py:183`); needs nx, ny >= 3"""
    assert not quadrilateral
    ny = nx if ny is None else ny
    assert nx >= 3 and ny >= 3, "periodic meshes need at least 3 cells per direction"
    cell_vert, cij = _square_cells(nx, ny, diagonal, periodic=True)
    cell_xy = cij * (np.array([L / nx, L / ny], dtype=np.float64))
    m = Mesh.from_cells(cell_vert, cell_xy=cell_xy, name=f"PeriodicSquareMesh({nx},{ny},L={L})")
    m.meta.update(nx=nx, ny=ny, diagonal=diagonal, periodic=True, L=L)
    return m


def UnitDiskMesh(refinement_level: int = 0) -> Mesh:
    """regularly refined polygonal approximation of the unit disk (`driver.py:185`).

    Starts from a regular octagon fan and splits every triangle into four per level, snapping new
    boundary vertices to the circle (the same construction Firedrake documents for UnitDiskMesh).
    """
    ang = np.arange(8) * (2 * np.pi / 8)
    verts = [(0.0, 0.0)] + [(np.cos(a), np.sin(a)) for a in ang]
    cells = [(0, 1 + i, 1 + (i + 1) % 8) for i in range(8)]
    verts = np.array(verts)
    cells = np.array(cells, dtype=np.int64)
    prolongations = []  # nested P1 spaces, used by the multigrid setup (finest first)
    for _ in range(refinement_level):
        nv = verts.shape[0]
        e = np.concatenate([cells[:, [1, 2]], cells[:, [2, 0]], cells[:, [0, 1]]])
        lo, hi = e.min(axis=1), e.max(axis=1)
        key = lo * nv + hi
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // nv, uniq % nv
        mid = 0.5 * (verts[a] + verts[b])
        on_bnd = (np.abs(np.hypot(*verts[a].T) - 1) < 1e-12) & (np.abs(np.hypot(*verts[b].T) - 1) < 1e-12)
        mid[on_bnd] /= np.hypot(*mid[on_bnd].T)[:, None]
        nc = cells.shape[0]
        m0, m1, m2 = nv + inv[:nc], nv + inv[nc:2 * nc], nv + inv[2 * nc:]
        v0, v1, v2 = cells.T
        cells = np.concatenate([
            np.stack([v0, m2, m1], 1), np.stack([v1, m0, m2], 1), np.stack([v2, m1, m0], 1), np.stack([m0, m1, m2], 1)
        ])
        verts = np.concatenate([verts, mid])
        import scipy.sparse as sp

        ne = uniq.size
        rows = np.concatenate([np.arange(nv), nv + np.arange(ne), nv + np.arange(ne)])
        cols = np.concatenate([np.arange(nv), a, b])
        vals = np.concatenate([np.ones(nv), np.full(ne, 0.5), np.full(ne, 0.5)])
        prolongations.insert(0, sp.coo_matrix((vals, (rows, cols)), shape=(nv + ne, nv)).tocsr())
    m = Mesh.from_cells(cells, vert_xy=verts, name=f"UnitDiskMesh({refinement_level})")
    m.meta.update(periodic=False, p1_prolongations=prolongations)
    return m


def RandomAffineCells(nc: int, seed: int = 123456789) -> Mesh:
    """`nc` independent triangles, each the unit right triangle under a random affine map.

    The condensation / back-substitution microbenchmark of BASELINE.json configs[4] (SURVEY.md §8d
    config 5): scale U[0.5,1.5], rotation U[0,2pi), shear U[-0.2,0.2]; no two cells are identical.
    Every facet is a boundary facet.
    """
    rng = np.random.default_rng(seed)
    sc = rng.uniform(0.5, 1.5, nc)
    th = rng.uniform(0.0, 2 * np.pi, nc)
    sh = rng.uniform(-0.2, 0.2, nc)
    ref = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    c, s = np.cos(th), np.sin(th)
    # A = scale * R(theta) * [[1, shear],[0, 1]]
    A = np.empty((nc, 2, 2))
    A[:, 0, 0] = sc * c
    A[:, 0, 1] = sc * (c * sh - s)
    A[:, 1, 0] = sc * s
    A[:, 1, 1] = sc * (s * sh + c)
    cell_xy = np.einsum("nij,vj->nvi", A, ref)
    cell_vert = np.arange(3 * nc, dtype=np.int64).reshape(nc, 3)
    m = Mesh.from_cells(cell_vert, cell_xy=cell_xy, name=f"RandomAffineCells({nc})")
    m.meta.update(periodic=False)
    return m
