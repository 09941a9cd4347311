"""CPU oracle for the passive-tracer path  --  TEST INFRASTRUCTURE ONLY (same rules as hdg_oracle.py:
only tests/, smoke() and bench.py's cpu_baseline leg may import it).  No Firedrake outputs to compare with; the
advection form is pinned against the reference's own `_tracer_advection` executed through oracle/miniufl.py
(tests/test_forms_golden.py); the CG projection is the standard L2 projection (`Function(V_CG).project(u)`).

Restates (reference file:line)

* ``_tracer_advection(chi, q, u, project_onto_cg=True)``      `timesteppers/common.py:110-129`
* the Chorin tracer update                                    `hdg_implicit.py:93-96,192-193`
* the IMEX tracer residuals                                   `hdg_imex.py:415-448,622-623,638-639`

independently of the engine: the CG_{k+1} space is found by *matching physical node coordinates*
(the engine numbers vertices / facets / interiors topologically), the CG mass matrix is assembled with
quadrature as a scipy sparse matrix and factorised (the engine applies it matrix-free inside a PCG), and
the advection form is integrated on the physical cell with the oracle's generic rules.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from incompressibleeulerhdg_b200 import refelem as R

__all__ = ["TracerOracle"]


class TracerOracle:
    def __init__(self, oracle):
        """`oracle` is an HDGOracle (geometry, tabulation and neighbour maps are reused)"""
        self.o = oracle
        self.d = oracle.k + 1
        self._cg_setup()

    # ------------------------------------------------------------------ CG_{k+1} by coordinate matching
    def _cg_setup(self):
        o, d = self.o, self.d
        nodes = R.lagrange_nodes_cell(d)
        xp = o.x0[:, None, :] + np.einsum("ncd,qd->nqc", o.J, nodes)  # [nc, nloc, 2]
        h = np.sqrt(o.detJ.min())
        key = np.round(xp.reshape(-1, 2) / (1e-6 * h)).astype(np.int64)
        _, inv = np.unique(key, axis=0, return_inverse=True)
        self.cellmap = inv.reshape(o.mesh.nc, -1)
        self.ndof = int(inv.max()) + 1
        # Lagrange basis at the quadrature points: L_j = sum_i Vinv[i, j] psi_i
        Vinv = R.nodal_to_modal_cell(d, nodes)
        self.L = np.einsum("ij,iq->jq", Vinv, o.phiQ)  # [nloc, nq]
        Mref = np.einsum("q,iq,jq->ij", o.wq, self.L, self.L)
        nloc = self.L.shape[0]
        rows = np.repeat(self.cellmap, nloc, axis=1).ravel()
        cols = np.tile(self.cellmap, (1, nloc)).ravel()
        vals = (o.detJ[:, None, None] * Mref[None]).ravel()
        self.M = sp.csr_matrix((vals, (rows, cols)), shape=(self.ndof, self.ndof))
        self.Mlu = spla.splu(self.M.tocsc())
        self.Vinv = Vinv

    def project_cg(self, Q):
        """`Function(V_CG).project(u)` (`common.py:121-122`), returned in the cell-wise modal basis"""
        o = self.o
        uq = o.eval_Q(Q)  # [nc, nq, 2]
        load = o.detJ[:, None, None] * np.einsum("q,nqc,jq->njc", o.wq, uq, self.L)
        b = np.zeros((self.ndof, 2))
        np.add.at(b, self.cellmap.ravel(), load.reshape(-1, 2))
        x = np.stack([self.Mlu.solve(b[:, c]) for c in range(2)], axis=1)
        xl = x[self.cellmap]  # [nc, nloc, 2]
        return np.einsum("ij,njc->nci", self.Vinv, xl)

    # ------------------------------------------------------------------ advection form
    def advection(self, q, Ucg):
        """M^-1 applied to  q div(chi u) dx - (chi+ - chi-)(un+ q+ - un- q-) dS   [nc, np]"""
        o = self.o
        qv = np.einsum("na,aq->nq", q, o.phiP)
        uv = o.eval_Q(Ucg)
        gpsi = np.einsum("ndc,iqd->niqc", o.Jinv, o.dphiQ)
        divu = np.einsum("nci,niqc->nq", Ucg, gpsi)
        gchi = np.einsum("ndc,aqd->naqc", o.Jinv, o.dphiP)
        integrand = np.einsum("nqc,naqc->naq", uv, gchi) + divu[:, None, :] * o.phiP[None]
        out = o.detJ[:, None] * np.einsum("q,nq,naq->na", o.wq, qv, integrand)
        # interior facets, seen from every cell: - int chi (max(un,0) q_K + min(un,0) q_nbr)
        uf = o.eval_Q_facet(Ucg)
        un = np.einsum("nec,neqc->neq", o.normal, uf)
        qf = np.einsum("na,eaq->neq", q, o.phiP_f)
        qn = o.nbr_facet_values(qf)
        flux = np.maximum(un, 0.0) * qf + np.minimum(un, 0.0) * qn
        flux = np.where((o.nbr >= 0)[:, :, None], flux, 0.0)
        out -= np.einsum("ne,q,neq,eaq->na", o.elen, o.wf, flux, o.phiP_f)
        return out / o.detJ[:, None]

    def total_mass(self, q):
        """int q dx"""
        o = self.o
        one = np.einsum("q,aq->a", o.wq, o.phiP)
        return float(np.einsum("n,na,a->", o.detJ, q, one))
