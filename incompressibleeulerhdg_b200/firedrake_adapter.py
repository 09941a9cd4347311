"""Firedrake <-> engine marshalling: the binding a reference maintainer selects with
``"pc_python_type": "incompressibleeulerhdg_b200.firedrake_adapter.SCPC"`` in place of
``"firedrake.SCPC"`` (`timesteppers/hdg_imex.py:132`).

Two layers:

* **Array level** (numpy only, unit-tested on CPU with fabricated layouts, `tests/test_firedrake_adapter.py`):
  :class:`NodalCellLayout` / :class:`NodalFacetLayout` convert between Firedrake's *nodal* coefficient
  arrays (``Function.dat.data`` indexed through ``cell_node_map().values``) and the engine's *modal* AoS
  arrays (``Q[nc,2,NQ1]``, ``p[nc,NP]``, ``lam[nf,k+1]``).  Nothing is assumed about FIAT's local node
  ordering or node variant (equispaced / GLL): the reference position of every local node
  is recovered from its *physical coordinates* and the cell's vertex coordinates, and one Vandermonde
  matrix is built per distinct node pattern.  Primal vectors transform with ``V^-1`` / ``V``, dual
  (residual) vectors with ``V^T`` / ``V^-T``.
* **Firedrake level** (needs ``import firedrake``; cannot be executed in this image):
  :class:`FiredrakeAdapter` pulls those arrays out of a mixed ``FunctionSpace`` and :class:`SCPC` is the
  PETSc Python-PC with the protocol the reference relies on: ``initialize/update/apply`` and the
  ``condensed_ksp.getIterationNumber()`` shim that `hdg_imex.py:265-271` reads.

The engine (``libhdg_b200.so``) is reached through :class:`engine.HDGEngine`, i.e. the C-ABI of
``include/hdg_b200.h``; there is no CPU fallback.
"""

from __future__ import annotations

import numpy as np

from . import refelem as R
from .mesh import Mesh

__all__ = ["NodalCellLayout", "NodalFacetLayout", "mesh_from_arrays", "FiredrakeAdapter", "SCPC"]


def mesh_from_arrays(cell_vert: np.ndarray, vert_xy: np.ndarray | None = None, cell_xy: np.ndarray | None = None) -> Mesh:
    """the engine's mesh from Firedrake's P1 coordinate field: ``cell_vert =
    mesh.coordinates.cell_node_map().values`` and ``vert_xy = mesh.coordinates.dat.data_ro`` (periodic
    meshes carry a DG coordinate field: pass the per-cell coordinates as ``cell_xy`` and the topological
    vertex ids of ``mesh.cell_closure`` as ``cell_vert``)"""
    return Mesh.from_cells(np.asarray(cell_vert), vert_xy=vert_xy, cell_xy=cell_xy, name="firedrake")


def _patterns(ref: np.ndarray, degree: int):
    """group entities by their pattern of reference node positions: (pattern nodes [npat, nloc, dim], id per entity)"""
    n = ref.shape[0]
    key = np.round(ref.reshape(n, -1) * (64.0 * max(degree, 1))).astype(np.int64)
    _, first, pid = np.unique(key, axis=0, return_index=True, return_inverse=True)
    return ref[first], pid.reshape(-1)


class NodalCellLayout:
    """nodal layout of a cell-wise space (DG_m, possibly vector valued) on the engine's mesh.

    cell_nodes [nc, nloc]  global node of every local node (Firedrake: ``V.cell_node_map().values``)
    node_xy    [nnodes, 2] physical node positions (Firedrake: coordinates interpolated into
                           ``VectorFunctionSpace(mesh, V.ufl_element().family(), m)``)
    Cells must be listed in the order of ``mesh`` (which keeps the caller's cell order)."""

    def __init__(self, mesh: Mesh, degree: int, cell_nodes: np.ndarray, node_xy: np.ndarray):
        self.mesh, self.degree = mesh, int(degree)
        self.cell_nodes = np.asarray(cell_nodes, dtype=np.int64)
        nc, nloc = self.cell_nodes.shape
        assert nc == mesh.nc and nloc == R.ncell(self.degree)
        x = mesh.cell_xy
        J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)  # [nc, c, d]
        d = np.asarray(node_xy, dtype=np.float64)[self.cell_nodes] - x[:, None, 0, :]
        if mesh.meta.get("periodic"):
            L = mesh.meta.get("L", 1.0)
            d = d - L * np.round(d / L)  # nodes of a wrapped cell may be stored on the far side
        ref = np.linalg.solve(J[:, None], d[..., None])[..., 0]  # xi = J^-1 (x - x0)  [nc, nloc, 2]
        assert ref.min() > -1e-8 and (ref.sum(axis=-1)).max() < 1 + 1e-8, "a node lies outside its cell"
        nodes, self.pid = _patterns(ref, self.degree)
        self.Vinv = np.array([R.nodal_to_modal_cell(self.degree, p) for p in nodes])  # modal = Vinv nodal
        self.V = np.array([np.linalg.inv(v) for v in self.Vinv])  # nodal = V modal, V[n, i] = phi_i(node n)

    # primal (coefficient) vectors -----------------------------------------------------------------
    def to_modal(self, data: np.ndarray) -> np.ndarray:
        """``Function.dat.data`` [nnodes] or [nnodes, ncomp] -> modal AoS [nc, ndof] or [nc, ncomp, ndof]"""
        loc = np.asarray(data)[self.cell_nodes]  # [nc, nloc(, ncomp)]
        Vi = self.Vinv[self.pid]
        if loc.ndim == 2:
            return np.einsum("nij,nj->ni", Vi, loc)
        return np.einsum("nij,njc->nci", Vi, loc)

    def from_modal(self, coef: np.ndarray, out: np.ndarray) -> np.ndarray:
        """modal AoS -> nodal values scattered into ``out`` (DG: every node belongs to one cell)"""
        V = self.V[self.pid]
        if coef.ndim == 2:
            out[self.cell_nodes] = np.einsum("nji,ni->nj", V, coef)
        else:
            out[self.cell_nodes] = np.einsum("nji,nci->njc", V, coef)
        return out

    # dual (residual) vectors: r(psi_i) = sum_j V[j, i] r(L_j) ----------------------------------------------
    def dual_to_modal(self, data: np.ndarray) -> np.ndarray:
        loc = np.asarray(data)[self.cell_nodes]
        V = self.V[self.pid]
        if loc.ndim == 2:
            return np.einsum("nji,nj->ni", V, loc)
        return np.einsum("nji,njc->nci", V, loc)

    def dual_from_modal(self, coef: np.ndarray, out: np.ndarray) -> np.ndarray:
        Vi = self.Vinv[self.pid]
        if coef.ndim == 2:
            out[self.cell_nodes] = np.einsum("nij,ni->nj", Vi, coef)
        else:
            out[self.cell_nodes] = np.einsum("nij,nci->njc", Vi, coef)
        return out


class NodalFacetLayout:
    """nodal layout of the trace space DGT_k: every facet owns k+1 nodes, shared by its two cells.

    cell_nodes [nc, 3 (k+1)]  ``V_trace.cell_node_map().values``; the k+1 nodes of one facet are assumed to be
                              consecutive local indices (FIAT numbers dofs entity by entity) -- *which* facet a
                              group belongs to and the order inside the group are recovered from coordinates
    node_xy    [nnodes, 2]    physical node positions (node sets symmetric about the facet midpoint)
    The engine's trace basis is Legendre on [0,1] along the facet's global direction (smaller -> larger
    topological vertex id, `mesh.py`)."""

    def __init__(self, mesh: Mesh, degree: int, cell_nodes: np.ndarray, node_xy: np.ndarray):
        self.mesh, self.degree = mesh, int(degree)
        k1 = self.degree + 1
        node_xy = np.asarray(node_xy, dtype=np.float64)
        nc, nf = mesh.nc, mesh.nf
        groups = np.asarray(cell_nodes, dtype=np.int64).reshape(nc, 3, k1)
        x = mesh.cell_xy
        a = x[:, [1, 2, 0]]  # start vertex of local facet e
        t = x[:, [2, 0, 1]] - a
        L = mesh.meta.get("L", 1.0)
        wrap = (lambda d: d - L * np.round(d / L)) if mesh.meta.get("periodic") else (lambda d: d)
        gx = a[:, 0:1, None, :] + wrap(node_xy[groups] - a[:, 0:1, None, :])  # [nc, g, j, 2], unwrapped near the cell
        cen = gx.mean(axis=2)  # centroid of every node group = midpoint of its facet
        mid = a + 0.5 * t
        dist = np.linalg.norm(cen[:, :, None, :] - mid[:, None, :, :], axis=-1)  # [nc, g, e]
        e_of_g = dist.argmin(axis=2)
        h = np.sqrt(np.abs(mesh.cell_area()))
        assert np.all(np.sort(e_of_g, axis=1) == np.arange(3)) and np.all(dist.min(axis=2) < 1e-8 * h[:, None]), \
            "could not match the trace node groups to the facets of their cell"
        cidx = np.arange(nc)[:, None]
        ag, tg = a[cidx, e_of_g], t[cidx, e_of_g]  # [nc, g, 2]
        s_loc = np.einsum("ngjc,ngc->ngj", gx - ag[:, :, None, :], tg) / np.einsum("ngc,ngc->ng", tg, tg)[..., None]
        flip = mesh.cell_flip[cidx, e_of_g].astype(bool)
        sg = np.where(flip[..., None], 1.0 - s_loc, s_loc)  # parameter along the global direction
        order = np.argsort(sg, axis=2)
        f = mesh.cell_facet[cidx, e_of_g]  # [nc, g]
        self.facet_nodes = np.full((nf, k1), -1, dtype=np.int64)
        self.s = np.zeros((nf, k1))
        self.facet_nodes[f] = np.take_along_axis(groups, order, axis=2)
        self.s[f] = np.take_along_axis(sg, order, axis=2)
        assert self.facet_nodes.min() >= 0
        # both cells of an interior facet must name the same nodes in the same global order
        assert np.array_equal(self.facet_nodes[f], np.take_along_axis(groups, order, axis=2)), \
            "the two cells of a facet disagree about its trace nodes"
        nodes, self.pid = _patterns(self.s[:, :, None], self.degree)
        self.Vinv = np.array([R.nodal_to_modal_facet(self.degree, p[:, 0]) for p in nodes])
        self.V = np.array([np.linalg.inv(v) for v in self.Vinv])

    def to_modal(self, data):
        return np.einsum("fij,fj->fi", self.Vinv[self.pid], np.asarray(data)[self.facet_nodes])

    def from_modal(self, coef, out):
        out[self.facet_nodes] = np.einsum("fji,fi->fj", self.V[self.pid], coef)
        return out

    def dual_to_modal(self, data):
        return np.einsum("fji,fj->fi", self.V[self.pid], np.asarray(data)[self.facet_nodes])

    def dual_from_modal(self, coef, out):
        out[self.facet_nodes] = np.einsum("fij,fi->fj", self.Vinv[self.pid], coef)
        return out


# ---------------------------------------------------------------------------------------------------
# Firedrake level (import-guarded; not executable in this image)
# ---------------------------------------------------------------------------------------------------
def _require_firedrake():
    try:
        import firedrake  # noqa: F401
    except ImportError as exc:  # pragma: no cover - Firedrake is absent here
        raise ImportError("incompressibleeulerhdg_b200.firedrake_adapter needs Firedrake for FiredrakeAdapter / SCPC; "
                          "the array-level layouts work without it") from exc
    return firedrake


class FiredrakeAdapter:
    """marshals the mixed space ``V = [DG_{k+1}]^2 x DG_k x DGT_k`` (`hdg_imex.py:65-69`) of a Firedrake
    mesh to the engine.  Serial (one rank) in this build; under MPI the DMPlex partition would be handed
    to ``partition.partition_mesh`` instead of ``strip_partition``."""

    def __init__(self, V, tau: float = 1.0, device: int = 0):  # pragma: no cover - needs Firedrake
        fd = _require_firedrake()
        from .engine import HDGEngine

        V_Q, V_p, V_tr = V.subfunctions if hasattr(V, "subfunctions") else V.split()
        mesh = V.mesh()
        coords = mesh.coordinates
        self.k = V_p.ufl_element().degree()
        self.mesh = mesh_from_arrays(coords.cell_node_map().values, vert_xy=coords.dat.data_ro)
        x = fd.SpatialCoordinate(mesh)

        def node_xy(space):
            W = fd.VectorFunctionSpace(mesh, space.ufl_element().family(), space.ufl_element().degree())
            return fd.Function(W).interpolate(x).dat.data_ro.copy()

        self.lay_Q = NodalCellLayout(self.mesh, self.k + 1, V_Q.cell_node_map().values, node_xy(V_Q))
        self.lay_p = NodalCellLayout(self.mesh, self.k, V_p.cell_node_map().values, node_xy(V_p))
        self.lay_l = NodalFacetLayout(self.mesh, self.k, V_tr.cell_node_map().values, node_xy(V_tr))
        self.sizes = (V_Q.dof_dset.size * 2, V_p.dof_dset.size, V_tr.dof_dset.size)
        self.engine = HDGEngine(self.mesh, self.k, tau=tau, device=device)
        self.engine.setup_poisson()
        try:
            self.engine.mg_setup()
        except ValueError:
            pass  # no nested P1 hierarchy for this mesh: facet-block-Jacobi CG

    def split(self, array):
        """mixed PETSc Vec array (field-major: Q interleaved by component, p, lambda) -> three views"""
        nQ, npp, nl = self.sizes
        return array[:nQ].reshape(-1, 2), array[nQ:nQ + npp], array[nQ + npp:nQ + npp + nl]

    def apply(self, x_array, y_array, rtol=1e-12, maxit=100000):
        """y = A^-1 x through the condensed engine solve; returns the trace-solve iteration count"""
        rQ, rp, rl = self.split(x_array)
        Q, p, l, its = self.engine.poisson_apply_host(self.lay_Q.dual_to_modal(rQ), self.lay_p.dual_to_modal(rp),
                                                      self.lay_l.dual_to_modal(rl), rtol=rtol, maxit=maxit, shift=False)
        yQ, yp, yl = self.split(y_array)
        self.lay_Q.from_modal(Q, yQ)
        self.lay_p.from_modal(p, yp)
        self.lay_l.from_modal(l, yl)
        return its


class _KSPShim:
    """what `hdg_imex.py:265-271` reads from the python context"""

    def __init__(self):
        self._its = 0

    def getIterationNumber(self):
        return self._its


def _pc_base():
    try:
        from firedrake import PCBase

        return PCBase
    except ImportError:
        return object


class SCPC(_pc_base()):
    """drop-in for ``firedrake.SCPC`` with ``pc_sc_eliminate_fields "0, 1"`` (`hdg_imex.py:129-170`): static
    condensation, trace Krylov solve with the GTMG-type preconditioner and back-substitution all run inside
    the engine; the PETSc options under ``condensed_field`` are therefore ignored (the engine's own CG
    tolerance is the reference's ``ksp_rtol`` 1e-12)."""

    needs_python_pmat = True

    def initialize(self, pc):  # pragma: no cover - needs Firedrake/PETSc
        _require_firedrake()
        _, P = pc.getOperators()
        ctx = P.getPythonContext()
        V = ctx.a.arguments()[0].function_space()
        self.adapter = FiredrakeAdapter(V)
        self.condensed_ksp = _KSPShim()

    def update(self, pc):  # the mixed-Poisson operator is constant
        pass

    def apply(self, pc, x, y):  # pragma: no cover - needs Firedrake/PETSc
        with x.getBuffer(readonly=True) as xa, y.getBuffer() as ya:
            self.condensed_ksp._its = self.adapter.apply(np.asarray(xa), np.asarray(ya))

    applyTranspose = apply  # symmetric up to the sign of the psi-row

    def view(self, pc, viewer=None):
        if viewer is not None:
            viewer.printfASCII("B200 HDG engine: static condensation + multigrid-preconditioned trace CG\n")
