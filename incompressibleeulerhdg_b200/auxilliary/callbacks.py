"""Callbacks invoked at the end of every timestep, with the reference's interface
(`src/auxilliary/callbacks.py:11-85`): ``callback(Q, p, t, q_tracer=None)`` and ``callback.reset()``.

``AnimationCallback`` saves velocity, pressure, vorticity (and the tracer) to a ParaView collection.
As in the reference (`callbacks.py:44-69`) the vorticity is the weak curl projected onto CG_{k+1}:
``tau xi dx == -eps : (grad tau (x) Q) dx + tau eps : (n (x) Q) ds`` with ``eps = [[0, 1], [-1, 0]]``
(:class:`VorticityProjector`, host-side scipy: output is off the hot path; the CG space is the one of the
tracer path, ``cgspace.build_cg_space``).  ``cell_vorticity_at_vertices`` is the strong cell-wise curl, kept
as a cheap diagnostic.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .. import refelem as R
from .vtk import VTKFile

__all__ = ["Callback", "AnimationCallback", "VorticityProjector", "cell_vorticity_at_vertices"]


class Callback(ABC):
    """Abstract base class"""

    @abstractmethod
    def __call__(self, Q, p, t, q_tracer=None):
        """invoke the callback for velocity/pressure (and tracer) fields at time t"""

    @abstractmethod
    def reset(self):
        """reset the callback"""


def cell_vorticity_at_vertices(Q):
    """curl Q = d_x Q_y - d_y Q_x of a cell-wise velocity at the three vertices of every cell [nc, 3]"""
    space = Q.function_space()
    mesh = space.mesh()
    coef = Q.to_host()  # [nc, 2, nQ1]
    dtab = R.dubiner_grad(space.degree, R.REF_VERTS)  # [nQ1, 3, 2] reference gradients
    x = mesh.cell_xy
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)  # J[n, c, d] = d x_c / d xi_d
    Jinv = np.linalg.inv(J)  # Jinv[n, d, c]
    gref = np.einsum("nci,ivd->ncvd", coef, dtab)  # d Q_c / d xi_d at the vertices
    grad = np.einsum("ncvd,nde->ncve", gref, Jinv)  # d Q_c / d x_e
    return grad[:, 1, :, 0] - grad[:, 0, :, 1]


class VorticityProjector:
    """``lvs, omega, Q_proxy = vorticity_solver(Q)`` of the reference (`callbacks.py:44-69`): the linear
    variational problem  tau xi dx == -eps:(grad tau (x) Q) dx + tau eps:(n (x) Q) ds  on CG_{degree},
    assembled once per (mesh, degree) and factorised with a sparse LU"""

    def __init__(self, mesh, degree: int):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        from ..cgspace import _local_entities, build_cg_space

        self.mesh, self.degree = mesh, int(degree)
        cg = build_cg_space(mesh, self.degree)
        self.cg = cg
        nloc, W = cg.nloc, cg.W
        x = mesh.cell_xy
        J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)
        self.detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        self.Jinv = np.linalg.inv(J)  # [n, d, c] = d xi_d / d x_c
        Mloc = W.T @ W  # int_T^ L_i L_j (orthonormal modal basis)
        rows = np.repeat(cg.cellmap, nloc, axis=1).ravel()
        cols = np.tile(cg.cellmap, (1, nloc)).ravel()
        M = sp.csr_matrix(((self.detJ[:, None, None] * Mloc[None]).ravel(), (rows, cols)), shape=(cg.ndof, cg.ndof))
        self._lu = spla.splu(M.tocsc())
        # tabulation: Lagrange basis L_j = sum_i W[i, j] psi_i and the velocity basis (same degree)
        xq, self.wq = R.triangle_quadrature_gj(2 * self.degree)
        self.psi = R.dubiner(self.degree, xq)  # [i, q]
        self.dL = np.einsum("ij,iqd->jqd", W, R.dubiner_grad(self.degree, xq))  # [j, q, d]
        s, self.wf = R.gauss_legendre(self.degree + 1)
        self.psi_f = np.array([R.dubiner(self.degree, R.facet_points(e, s)) for e in range(3)])  # [e, i, q]
        self.L_f = np.einsum("ij,eiq->ejq", W, self.psi_f)
        va, vb = x[:, [1, 2, 0]], x[:, [2, 0, 1]]
        t = vb - va
        self.elen = np.hypot(t[..., 0], t[..., 1])
        self.normal = np.stack([t[..., 1], -t[..., 0]], axis=-1) / self.elen[..., None]
        self.on_boundary = mesh.facet_cell[mesh.cell_facet, 1] < 0  # [nc, 3]
        ents, _ = _local_entities(self.degree)
        self.vertex_nodes = [next(j for j, en in enumerate(ents) if en == ("v", v)) for v in range(3)]

    def __call__(self, Qcoef: np.ndarray) -> np.ndarray:
        """nodal values of the projected vorticity [ndof] for modal velocity coefficients [nc, 2, ndof_loc]"""
        Qq = np.einsum("nci,iq->nqc", Qcoef, self.psi)  # velocity at the cell points
        gL = np.einsum("ndc,jqd->njqc", self.Jinv, self.dL)  # physical gradient of L_j
        integrand = gL[..., 0] * Qq[:, None, :, 1] - gL[..., 1] * Qq[:, None, :, 0]  # eps : (grad tau (x) Q)
        b = -self.detJ[:, None] * np.einsum("q,njq->nj", self.wq, integrand)
        Qf = np.einsum("nci,eiq->neqc", Qcoef, self.psi_f)
        nxq = self.normal[:, :, None, 0] * Qf[..., 1] - self.normal[:, :, None, 1] * Qf[..., 0]  # eps : (n (x) Q)
        nxq = np.where(self.on_boundary[:, :, None], nxq, 0.0)
        b += np.einsum("ne,q,neq,ejq->nj", self.elen, self.wf, nxq, self.L_f)
        rhs = np.bincount(self.cg.cellmap.ravel(), weights=b.ravel(), minlength=self.cg.ndof)
        return self._lu.solve(rhs)

    def at_vertices(self, Qcoef: np.ndarray) -> np.ndarray:
        """the projected vorticity at the three vertices of every cell [nc, 3] (what the VTK writer samples)"""
        omega = self(Qcoef)
        return omega[self.cg.cellmap[:, self.vertex_nodes]]


class AnimationCallback(Callback):
    """Save fields to disk (`callbacks.py:28-85`)"""

    def __init__(self, filename):
        self.filename = filename
        self.reset()

    def reset(self):
        """re-open the file"""
        self.outfile = VTKFile(self.filename, mode="w")
        self._projector = None

    def vorticity_solver(self, Q):
        """cached per velocity space, like the reference's ``functools.cache`` (`callbacks.py:43`)"""
        space = Q.function_space()
        if self._projector is None or self._projector.mesh is not space.mesh():
            self._projector = VorticityProjector(space.mesh(), space.degree)
        return self._projector

    def __call__(self, Q, p, t, q_tracer=None):
        fields = [Q, p, ("vorticity", self.vorticity_solver(Q).at_vertices(Q.to_host()))]
        if q_tracer is not None:
            fields.append(q_tracer)
        self.outfile.write(*fields, time=t)
