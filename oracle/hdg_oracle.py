"""CPU oracle for the HDG hot path of eikehmueller/IncompressibleEulerHDG  --  TEST INFRASTRUCTURE ONLY.

PARITY: the reference ships no tests, golden vectors or recorded outputs and its
arithmetic lives in Firedrake/Slate/PETSc/MUMPS, none of which is installed here or pinned by the
reference (`requirements.txt:1-3`), so OUTPUTS of the reference are unavailable ("parity unpinned" in
that sense).  This file *restates* the reference's UFL forms with plain numpy quadrature.  The forms
themselves ARE pinned: `oracle/miniufl.py` executes the reference's own form-building code (cut out of
`hdg_imex.py` / `hdg_implicit.py` / `common.py` with ast) and `tests/test_forms_golden.py` checks every
operator and right-hand side assembled here against it to round-off.  Further anchors: (i) the analytic
Taylor-Green solution of `model_problems.py:56-105`, (ii) structural invariants (symmetry, null vector
(0,1,1), condensed == monolithic), and (iii) observed convergence rates.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this
module.  The product path (``incompressibleeulerhdg_b200``) never does.

What is restated (reference file:line):

* local mixed-Poisson operator      `hdg_imex.py:123-127`, `_pressure_gradient :333-340`,
                                    `_Gamma :342-351`; written out again `hdg_implicit.py:133-143`
* static condensation (Slate SCPC)  `hdg_imex.py:128-133`  S_K = D - C A^-1 B by dense LU
* trace solve                       `hdg_imex.py:134-137`  (here: sparse direct, constants pinned)
* back-substitution                 SCPC.apply
* weak divergence RHS               `hdg_imex.py:353-365`, Chorin RHS `hdg_implicit.py:145`
* pressure shift / null space       `hdg_imex.py:471-489`
* trace reconstruction              `hdg_imex.py:450-469`
* BDM projection                    `common.py:59-70,91-108`
* f_impl (advection+penalty)        `hdg_imex.py:313-331`; Chorin operator `hdg_implicit.py:103-125`
* 1/h_F                             `common.py:36-57`

Conventions: FP64; coefficient arrays are AoS: Q[nc, 2, nQ1], p[nc, np], lam[nf, k+1] in the modal
bases of ``incompressibleeulerhdg_b200.refelem`` (Dubiner on cells, Legendre on facets along the
global facet direction).  The local matrices are integrated on the *physical* cell with generic
quadrature and solved with dense LU -- deliberately not the closed-form reference-tensor
factorisation the CUDA engine uses, so that the two derivations check each other.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from incompressibleeulerhdg_b200 import refelem as R

__all__ = ["HDGOracle"]


class HDGOracle:
    def __init__(self, mesh, k: int, tau: float = 1.0, alpha_penalty: float = 1.0, flux: str = "upwind",
                 nq_facet: int | None = None):
        self.mesh = mesh
        self.k = k
        self.tau = float(tau)
        self.alpha = float(alpha_penalty)
        self.flux = flux
        self.nQ1 = R.ncell(k + 1)
        self.nQ = 2 * self.nQ1
        self.np_ = R.ncell(k)
        self.nl1 = k + 1
        self.nl = 3 * (k + 1)
        self.nA = self.nQ + self.np_
        # facet quadrature: TSFC estimates degree 3k+3 for the upwind term
        self.nq_facet = nq_facet if nq_facet is not None else (3 * k + 4 + 1) // 2
        self._tabulate()
        self._geometry()

    # ------------------------------------------------------------------ reference tabulation
    def _tabulate(self):
        k = self.k
        self.xq, self.wq = R.triangle_quadrature(3 * k + 4)
        self.phiQ = R.dubiner(k + 1, self.xq)  # [nQ1, nq]
        self.dphiQ = R.dubiner_grad(k + 1, self.xq)  # [nQ1, nq, 2]
        self.phiP = R.dubiner(k, self.xq)
        self.dphiP = R.dubiner_grad(k, self.xq)
        self.sq, self.wf = R.gauss_legendre(self.nq_facet)
        self.phiQ_f = np.array([R.dubiner(k + 1, R.facet_points(e, self.sq)) for e in range(3)])  # [3,nQ1,nqf]
        self.dphiQ_f = np.array([R.dubiner_grad(k + 1, R.facet_points(e, self.sq)) for e in range(3)])
        self.phiP_f = np.array([R.dubiner(k, R.facet_points(e, self.sq)) for e in range(3)])  # [3,np,nqf]
        # trace basis along the global direction: s_glob = s (flip 0) or 1-s (flip 1)
        self.ell = np.array([R.legendre01(k, self.sq), R.legendre01(k, 1.0 - self.sq)])  # [2,k+1,nqf]

    # ------------------------------------------------------------------ per-cell geometry
    def _geometry(self):
        x = self.mesh.cell_xy
        J = np.empty((self.mesh.nc, 2, 2))
        J[:, :, 0] = x[:, 1] - x[:, 0]
        J[:, :, 1] = x[:, 2] - x[:, 0]
        self.detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        assert np.all(self.detJ > 0)
        self.Jinv = np.linalg.inv(J)  # Jinv[d, c] = d xi_d / d x_c
        va = x[:, [1, 2, 0]]
        vb = x[:, [2, 0, 1]]
        t = vb - va  # [nc, 3, 2]
        self.elen = np.hypot(t[..., 0], t[..., 1])
        self.normal = np.stack([t[..., 1], -t[..., 0]], axis=-1) / self.elen[..., None]  # outward (CCW cells)
        self.x0 = x[:, 0]
        self.J = J
        # 1/h_F on facets (`common.py:36-57`)
        self.hF_inv = 1.0 / self.mesh.facet_length()
        fc = self.mesh.facet_cell
        self.interior = fc[:, 1] >= 0
        # neighbour across local facet e, and the neighbour's local facet index
        nc = self.mesh.nc
        self.nbr = np.full((nc, 3), -1, dtype=np.int64)
        self.nbr_e = np.full((nc, 3), -1, dtype=np.int64)
        fl = self.mesh.facet_local
        f_int = np.nonzero(self.interior)[0]
        c0, c1 = fc[f_int, 0], fc[f_int, 1]
        e0, e1 = fl[f_int, 0], fl[f_int, 1]
        self.nbr[c0, e0] = c1
        self.nbr_e[c0, e0] = e1
        self.nbr[c1, e1] = c0
        self.nbr_e[c1, e1] = e0

    def phys_points(self):
        """physical coordinates of the cell quadrature points [nc, nq, 2]"""
        return self.x0[:, None, :] + np.einsum("ncd,qd->nqc", self.J, self.xq)

    def phys_facet_points(self):
        """physical coordinates of the facet quadrature points [nc, 3, nqf, 2]"""
        ref = np.array([R.facet_points(e, self.sq) for e in range(3)])  # [3,nqf,2]
        return self.x0[:, None, None, :] + np.einsum("ncd,eqd->neqc", self.J, ref)

    # ------------------------------------------------------------------ local operators
    def local_blocks(self):
        """dense local blocks of a_mixed_poisson (`hdg_imex.py:123-127`) for every cell.

        Returns dict with M [nc,nQ,nQ], B [nc,np,nQ], E [nc,nl,nQ], F [nc,nl,np], T [nc,np,np],
        G [nc,nl,nl] such that, with rows (w,psi,mu) and columns (u,phi,lambda):

            [  M   -B^T    E^T  ]
            [  B     T   -tau F^T ]
            [  E   tau F  -tau G ]
        """
        nc = self.mesh.nc
        nQ1, np_, nl1 = self.nQ1, self.np_, self.nl1
        w = self.wq
        detJ = self.detJ
        # physical gradients of velocity basis: gphi[n, i, q, c] = sum_d Jinv[n,d,c] dphi[i,q,d]
        gphi = np.einsum("ndc,iqd->niqc", self.Jinv, self.dphiQ)
        M1 = np.einsum("q,iq,jq->ij", w, self.phiQ, self.phiQ)
        M = np.zeros((nc, 2 * nQ1, 2 * nQ1))
        for c in range(2):
            M[:, c * nQ1:(c + 1) * nQ1, c * nQ1:(c + 1) * nQ1] = detJ[:, None, None] * M1
        B = np.zeros((nc, np_, 2 * nQ1))
        for c in range(2):
            B[:, :, c * nQ1:(c + 1) * nQ1] = detJ[:, None, None] * np.einsum("q,aq,niq->nai", w, self.phiP, gphi[..., c])
        E = np.zeros((nc, 3 * nl1, 2 * nQ1))
        F = np.zeros((nc, 3 * nl1, np_))
        T = np.zeros((nc, np_, np_))
        G = np.zeros((nc, 3 * nl1, 3 * nl1))
        flip = self.mesh.cell_flip
        for e in range(3):
            ell = self.ell[flip[:, e]]  # [nc, k+1, nqf]
            le = self.elen[:, e]
            rows = slice(e * nl1, (e + 1) * nl1)
            for c in range(2):
                E[:, rows, c * nQ1:(c + 1) * nQ1] = (le * self.normal[:, e, c])[:, None, None] * np.einsum(
                    "q,nmq,iq->nmi", self.wf, ell, self.phiQ_f[e])
            F[:, rows, :] = le[:, None, None] * np.einsum("q,nmq,aq->nma", self.wf, ell, self.phiP_f[e])
            T += self.tau * le[:, None, None] * np.einsum("q,aq,bq->ab", self.wf, self.phiP_f[e], self.phiP_f[e])
            G[:, rows, rows] = le[:, None, None] * np.einsum("q,nmq,nlq->nml", self.wf, ell, ell)
        return dict(M=M, B=B, E=E, F=F, T=T, G=G)

    def local_system(self):
        """A_K [nc,nA,nA], B_K [nc,nA,nl], C_K [nc,nl,nA], D_K [nc,nl,nl]"""
        b = self.local_blocks()
        nc, nQ, np_, nl = self.mesh.nc, self.nQ, self.np_, self.nl
        A = np.zeros((nc, self.nA, self.nA))
        A[:, :nQ, :nQ] = b["M"]
        A[:, :nQ, nQ:] = -np.swapaxes(b["B"], 1, 2)
        A[:, nQ:, :nQ] = b["B"]
        A[:, nQ:, nQ:] = b["T"]
        Bk = np.zeros((nc, self.nA, nl))
        Bk[:, :nQ, :] = np.swapaxes(b["E"], 1, 2)
        Bk[:, nQ:, :] = -self.tau * np.swapaxes(b["F"], 1, 2)
        Ck = np.zeros((nc, nl, self.nA))
        Ck[:, :, :nQ] = b["E"]
        Ck[:, :, nQ:] = self.tau * b["F"]
        Dk = -self.tau * b["G"]
        return A, Bk, Ck, Dk

    # ------------------------------------------------------------------ dof numbering (monolithic)
    def trace_dofs(self):
        """global trace dof numbers per cell [nc, nl]"""
        return (self.mesh.cell_facet.astype(np.int64)[:, :, None] * self.nl1 + np.arange(self.nl1)[None, None, :]).reshape(
            self.mesh.nc, self.nl)

    def condensed_local(self):
        """S_K = D - C A^-1 B per cell by dense LU with partial pivoting (what Slate/Eigen does)"""
        A, Bk, Ck, Dk = self.local_system()
        AinvB = np.linalg.solve(A, Bk)
        return Dk - Ck @ AinvB

    def assemble_trace_matrix(self, SK=None):
        """global S = sum_K P_K^T S_K P_K as CSR (`hdg_imex.py:135`, mat_type aij)"""
        SK = self.condensed_local() if SK is None else SK
        td = self.trace_dofs()
        n = self.mesh.nf * self.nl1
        rows = np.repeat(td[:, :, None], self.nl, axis=2).ravel()
        cols = np.repeat(td[:, None, :], self.nl, axis=1).ravel()
        return sp.csr_matrix((SK.ravel(), (rows, cols)), shape=(n, n))

    def assemble_monolithic(self):
        """the full (u,phi,lambda) sparse operator, field-major numbering"""
        A, Bk, Ck, Dk = self.local_system()
        nc, nA, nl = self.mesh.nc, self.nA, self.nl
        nQ, np_ = self.nQ, self.np_
        offp = nc * nQ
        offl = offp + nc * np_
        cd = np.empty((nc, nA), dtype=np.int64)
        cd[:, :nQ] = np.arange(nc)[:, None] * nQ + np.arange(nQ)[None, :]
        cd[:, nQ:] = offp + np.arange(nc)[:, None] * np_ + np.arange(np_)[None, :]
        td = offl + self.trace_dofs()
        N = offl + self.mesh.nf * self.nl1

        def coo(blk, r, c):
            rr = np.repeat(r[:, :, None], c.shape[1], axis=2).ravel()
            cc = np.repeat(c[:, None, :], r.shape[1], axis=1).ravel()
            return sp.coo_matrix((blk.ravel(), (rr, cc)), shape=(N, N))

        K = coo(A, cd, cd) + coo(Bk, cd, td) + coo(Ck, td, cd) + coo(Dk, td, td)
        return K.tocsr(), (offp, offl, N)

    # ------------------------------------------------------------------ null space / shifts
    def null_vector_trace(self):
        """coefficient vector of lambda == 1 (Legendre mode 0 equals 1 on [0,1])"""
        z = np.zeros((self.mesh.nf, self.nl1))
        z[:, 0] = 1.0
        return z

    def const_p(self):
        """coefficient vector of p == 1: Dubiner mode 0 is the constant sqrt(2)"""
        z = np.zeros((self.mesh.nc, self.np_))
        z[:, 0] = 1.0 / np.sqrt(2.0)
        return z

    def integral_p(self, p):
        """int_Omega p dx (`hdg_imex.py:476`)"""
        return float(np.sum(self.detJ * p[:, 0]) / np.sqrt(2.0))

    def shift_pressure(self, p, lam):
        """`_shift_pressure` `hdg_imex.py:471-478`"""
        shift = self.integral_p(p) / self.mesh.volume
        return p - shift * self.const_p(), lam - shift * self.null_vector_trace()

    # ------------------------------------------------------------------ solves
    def consistency_defect(self, Ru, Rp, Rl):
        """y^T R for the left null vector y = (0, -1, 1) of the mixed operator (rows w, psi, mu).

        Equals z^T r for the condensed right-hand side r.  Zero for every right-hand side built
        from `_weak_divergence` (`hdg_imex.py:353-365`); non-zero in general for the Chorin
        right-hand side of `hdg_implicit.py:145`, whose broken divergence does not integrate to 0.
        """
        return float(Rl[:, 0].sum() - np.sum(Rp * self.const_p()))

    def _project_trace_rhs(self, Ru, Rp, Rl):
        """make the singular system consistent by removing the defect from the trace row,
        R_lambda <- R_lambda - z (y^T R)/(z^T z).  This is the engine's documented semantics for
        inconsistent data (an orthogonal projection of the condensed rhs onto range(S))."""
        d = self.consistency_defect(Ru, Rp, Rl)
        return Rl - self.null_vector_trace() * (d / self.mesh.nf)

    def solve_monolithic(self, Ru, Rp, Rl):
        """sparse-direct solve of the full system with one trace dof pinned.

        This is the route `hdg_implicit.py:146` takes (default direct LU).  Returns (u, phi, lam)
        *after* `_shift_pressure` so the result is unique.
        """
        K, (offp, offl, N) = self.assemble_monolithic()
        Rl = self._project_trace_rhs(Ru, Rp, Rl)
        rhs = np.concatenate([Ru.ravel(), Rp.ravel(), Rl.ravel()])
        keep = np.ones(N, dtype=bool)
        keep[offl] = False
        Kr = K[keep][:, keep].tocsc()
        x = np.zeros(N)
        x[keep] = spla.splu(Kr).solve(rhs[keep])
        u = x[:offp].reshape(self.mesh.nc, 2, self.nQ1)
        p = x[offp:offl].reshape(self.mesh.nc, self.np_)
        lam = x[offl:N].reshape(self.mesh.nf, self.nl1)
        p, lam = self.shift_pressure(p, lam)
        return u, p, lam

    def solve_condensed(self, Ru, Rp, Rl, return_parts=False):
        """static condensation route (`hdg_imex.py:128-137`): forward elimination, trace solve,
        back-substitution, all with dense local LU like Slate.  Result after `_shift_pressure`."""
        A, Bk, Ck, Dk = self.local_system()
        nc = self.mesh.nc
        Rloc = np.concatenate([Ru.reshape(nc, self.nQ), Rp.reshape(nc, self.np_)], axis=1)
        x0 = np.linalg.solve(A, Rloc[:, :, None])[:, :, 0]
        SK = Dk - Ck @ np.linalg.solve(A, Bk)
        td = self.trace_dofs()
        n = self.mesh.nf * self.nl1
        r = Rl.ravel().copy()
        np.subtract.at(r, td.ravel(), np.einsum("nla,na->nl", Ck, x0).ravel())
        z = self.null_vector_trace().ravel()
        r -= z * (z @ r) / (z @ z)
        S = self.assemble_trace_matrix(SK)
        lam = np.zeros(n)
        lam[1:] = spla.splu(S[1:][:, 1:].tocsc()).solve(r[1:])
        xl = np.linalg.solve(A, (Rloc - np.einsum("nal,nl->na", Bk, lam[td]))[:, :, None])[:, :, 0]
        u = xl[:, :self.nQ].reshape(nc, 2, self.nQ1)
        p = xl[:, self.nQ:]
        lam = lam.reshape(self.mesh.nf, self.nl1)
        p, lam = self.shift_pressure(p, lam)
        if return_parts:
            return u, p, lam, dict(S=S, r=r.reshape(self.mesh.nf, self.nl1), SK=SK)
        return u, p, lam

    # ------------------------------------------------------------------ evaluation helpers
    def eval_Q(self, Q):
        """velocity at cell quadrature points [nc, nq, 2]"""
        return np.einsum("nci,iq->nqc", Q, self.phiQ)

    def eval_Q_facet(self, Q):
        """velocity at the facet quadrature points of each cell [nc, 3, nqf, 2] (cell-local parameter)"""
        return np.einsum("nci,eiq->neqc", Q, self.phiQ_f)

    def nbr_facet_values(self, vals):
        """values of the neighbouring cell at the *same physical* facet points.

        `vals` [nc,3,nqf,...] in cell-local facet parametrisation; the neighbour traverses the facet
        in the opposite direction, so its points are reversed.  Boundary facets return zeros.
        """
        out = np.zeros_like(vals)
        has = self.nbr >= 0
        c, e = np.nonzero(has)
        out[c, e] = vals[self.nbr[c, e], self.nbr_e[c, e]][:, ::-1]
        return out

    def project_cell(self, fun, space="Q"):
        """L2 projection of a callable f(x,y) onto the modal cell basis (orthonormal => quadrature)"""
        xp = self.phys_points()
        v = fun(xp[..., 0], xp[..., 1])
        if space == "Q":
            v = np.stack(np.broadcast_arrays(*v), axis=-1)  # [nc,nq,2]
            return np.einsum("q,nqc,iq->nci", self.wq, v, self.phiQ)
        return np.einsum("q,nq,aq->na", self.wq, np.broadcast_to(v, xp.shape[:2]), self.phiP)

    def interpolate_cell(self, fun, space="Q", nodes=None):
        """nodal interpolation at (equispaced) Lagrange nodes followed by nodal->modal conversion,
        the analogue of Firedrake's `Function.interpolate` (`hdg_imex.py:520-521`)"""
        m = self.k + 1 if space == "Q" else self.k
        nodes = R.lagrange_nodes_cell(m) if nodes is None else nodes
        Vinv = R.nodal_to_modal_cell(m, nodes)
        xp = self.x0[:, None, :] + np.einsum("ncd,qd->nqc", self.J, nodes)
        v = fun(xp[..., 0], xp[..., 1])
        if space == "Q":
            v = np.stack(np.broadcast_arrays(*v), axis=-1)
            return np.einsum("iq,nqc->nci", Vinv, v)
        return np.einsum("aq,nq->na", Vinv, np.broadcast_to(v, xp.shape[:2]))

    def l2_error_Q(self, Q, fun):
        xp = self.phys_points()
        ex = np.stack(np.broadcast_arrays(*fun(xp[..., 0], xp[..., 1])), axis=-1)
        d = self.eval_Q(Q) - ex
        return float(np.sqrt(np.einsum("n,q,nqc->", self.detJ, self.wq, d * d)))

    def l2_error_p(self, p, fun):
        xp = self.phys_points()
        d = np.einsum("na,aq->nq", p, self.phiP) - fun(xp[..., 0], xp[..., 1])
        return float(np.sqrt(np.einsum("n,q,nq->", self.detJ, self.wq, d * d)))

    def l2_norm_Q(self, Q):
        return float(np.sqrt(np.einsum("n,nci->", self.detJ, Q * Q)))

    # ------------------------------------------------------------------ right-hand sides
    def mass_Q(self, Q):
        """(w, Q) dx  as a dual vector [nc,2,nQ1] (quadrature, not the M = detJ I shortcut)"""
        v = self.eval_Q(Q)
        return self.detJ[:, None, None] * np.einsum("q,nqc,iq->nci", self.wq, v, self.phiQ)

    def cell_divergence(self, Q):
        """int_K psi div Q dx  [nc, np]   (Chorin RHS kernel, `hdg_implicit.py:145`)"""
        gphi = np.einsum("ndc,iqd->niqc", self.Jinv, self.dphiQ)
        div = np.einsum("nci,niqc->nq", Q, gphi)
        return self.detJ[:, None] * np.einsum("q,nq,aq->na", self.wq, div, self.phiP)

    def weak_divergence(self, Q):
        """`_weak_divergence(psi, Q)` `hdg_imex.py:353-365` as a dual vector [nc, np]:

        int_K psi div Q - 1/2 int_{dK int} psi n.(Q_K - Q_nbr) - int_{dK bnd} psi n.Q
        """
        out = self.cell_divergence(Q)
        Qf = self.eval_Q_facet(Q)
        Qn = self.nbr_facet_values(Qf)
        interior = (self.nbr >= 0)[:, :, None, None]
        jump = np.where(interior, 0.5 * (Qf - Qn), Qf)
        nj = np.einsum("nec,neqc->neq", self.normal, jump)
        out -= np.einsum("ne,q,neq,eaq->na", self.elen, self.wf, nj, self.phiP_f)
        return out

    def weak_divergence_fun(self, Xq, Xf):
        """`_weak_divergence(psi, X)` for a general piecewise-smooth X given by its values at the
        cell quadrature points Xq [nc,nq,2] and the facet points Xf [nc,3,nqf,2] (cell-local
        parametrisation).  Integrated by parts (an identity for the piecewise polynomials used,
        `hdg_imex.py:204-207`):   - int_K grad psi . X + int_{dK int} psi n.avg(X)
        """
        gpsi = np.einsum("ndc,aqd->naqc", self.Jinv, self.dphiP)
        out = -self.detJ[:, None] * np.einsum("q,nqc,naqc->na", self.wq, Xq, gpsi)
        Xn = self.nbr_facet_values(Xf)
        interior = (self.nbr >= 0)[:, :, None, None]
        av = np.where(interior, 0.5 * (Xf + Xn), 0.0)
        nav = np.einsum("nec,neqc->neq", self.normal, av)
        out += np.einsum("ne,q,neq,eaq->na", self.elen, self.wf, nav, self.phiP_f)
        return out

    def pressure_gradient(self, p, lam):
        """g(w, p, lambda) `hdg_imex.py:333-340` as a dual vector [nc,2,nQ1]"""
        b = self.local_blocks()
        td = self.trace_dofs()
        lamK = lam.ravel()[td]
        g = np.einsum("nai,na->ni", b["B"], p) - np.einsum("nli,nl->ni", b["E"], lamK)
        return g.reshape(self.mesh.nc, 2, self.nQ1)

    def reconstruct_trace(self, Q, p):
        """`_reconstruct_trace` `hdg_imex.py:450-469`: facet-wise L2 projection of the average of
        (Q.n_K / tau + p) over the adjacent cells"""
        Qf = self.eval_Q_facet(Q)
        pf = np.einsum("na,eaq->neq", p, self.phiP_f)
        val = np.einsum("nec,neqc->neq", self.normal, Qf) + self.tau * pf
        flip = self.mesh.cell_flip
        rhs = np.zeros((self.mesh.nf, self.nl1))
        for e in range(3):
            ell = self.ell[flip[:, e]]
            contrib = self.elen[:, e, None] * np.einsum("q,nq,nmq->nm", self.wf, val[:, e], ell)
            np.add.at(rhs, self.mesh.cell_facet[:, e], contrib)
        mult = np.where(self.interior, 2.0, 1.0)
        return rhs / (self.tau * mult * self.mesh.facet_length())[:, None]

    # ------------------------------------------------------------------ BDM projection
    def _bdm_setup(self):
        """Local BDM_{k+1} degrees of freedom as linear functionals on [P_{k+1}]^2 (reference
        element, Piola-mapped), following the classical definition used by FIAT's BDM element:

        * facet moments   int_F (v.n) q ds           for q in P_{k+1}(F)           3(k+2)
        * interior moments int_K v . w dx            for w in the Nedelec space of
                                                     the first kind NED1_k(K)      k(k+2)

        `project_bdm` (`common.py:91-108`) is basis independent: average the facet
        moments between the two cells, zero them on the boundary, keep the interior moments.
        """
        if hasattr(self, "_bdm"):
            return self._bdm
        k1 = self.k + 1  # BDM degree
        nQ1 = self.nQ1
        # facet moment functionals, cell-local parametrisation, physical normal component:
        #   m_{e,j}(v) = int_0^1 (v.n)(s) l_j(s) ds        (length factor dropped: same on both sides)
        sq, wf = R.gauss_legendre(k1 + 2)
        legs = R.legendre01(k1, sq)  # [k1+1, nq]
        phf = np.array([R.dubiner(k1, R.facet_points(e, sq)) for e in range(3)])  # [3,nQ1,nq]
        self._bdm_facet = np.einsum("q,jq,eiq->eji", wf, legs, phf)  # [3, k1+1, nQ1]
        # interior functionals: v -> int_K^ (J^-1 v) . w^ dxi, w^ in NED1_k(K^) = [P_{k-1}]^2 + S_k
        xq, wq = R.triangle_quadrature(2 * k1 + 2)
        ph = R.dubiner(k1, xq)
        ned = _nedelec1_basis(self.k, xq)  # [nint, nq, 2]
        self._bdm_int = np.einsum("q,wqd,iq->wdi", wq, ned, ph)  # [nint, 2(d), nQ1]
        self._bdm = True
        return True

    def _bdm_dof_matrix(self):
        """L [nc, nQ, nQ]: all BDM functionals applied to the modal velocity basis, per cell"""
        self._bdm_setup()
        nc, nQ1 = self.mesh.nc, self.nQ1
        k1 = self.k + 1
        nfm = 3 * (k1 + 1)
        nint = self._bdm_int.shape[0]
        assert nfm + nint == self.nQ, (nfm, nint, self.nQ)
        L = np.zeros((nc, self.nQ, 2, nQ1))
        for e in range(3):
            for c in range(2):
                L[:, e * (k1 + 1):(e + 1) * (k1 + 1), c, :] = self.normal[:, e, c][:, None, None] * self._bdm_facet[e][None]
        # interior: (J^-1 v)_d = sum_c Jinv[d,c] v_c
        L[:, nfm:, :, :] = np.einsum("ndc,wdi->nwci", self.Jinv, self._bdm_int)
        return L.reshape(nc, self.nQ, self.nQ), nfm

    def project_bdm(self, Q):
        """`project_bdm` `common.py:91-108` in the cell-wise [P_{k+1}]^2 representation"""
        L, nfm = self._bdm_dof_matrix()
        nc = self.mesh.nc
        k1 = self.k + 1
        mom = np.einsum("nri,ni->nr", L, Q.reshape(nc, self.nQ))
        fm = mom[:, :nfm].reshape(nc, 3, k1 + 1)
        # neighbour moment of the same functional: its parametrisation is reversed and its normal
        # is opposite:  l_j(1-s) = (-1)^j l_j(s),  n_nbr = -n
        sign = -((-1.0) ** np.arange(k1 + 1))
        has = self.nbr >= 0
        c, e = np.nonzero(has)
        avg = np.zeros_like(fm)  # boundary facets: DirichletBC zero (`common.py:106-107`)
        avg[c, e] = 0.5 * (fm[c, e] + sign[None, :] * fm[self.nbr[c, e], self.nbr_e[c, e]])
        mom_new = mom.copy()
        mom_new[:, :nfm] = avg.reshape(nc, nfm)
        return np.linalg.solve(L, mom_new[:, :, None])[:, :, 0].reshape(nc, 2, self.nQ1)

    # ------------------------------------------------------------------ f_impl
    def f_impl_apply(self, Q, Qstar):
        """f^{im}(w, Q; Q*) `hdg_imex.py:313-331` as a dual vector [nc,2,nQ1].

        Per cell K with outward normal n, s = Q*.n (single valued for BDM Q*), nbr = neighbour:
          - int_K w_c (Q* . grad) Q_c
          + 1/2 int_{dK int} s (Q_K - Q_nbr).w            ['+' / '-' symmetric form of the dS term]
          - alpha/h_F int_{dK int} ((Q_K - Q_nbr).n)(w.n)  [4 avg(hF^-1) avg(Q.n) avg(w.n)]
          - alpha/h_F int_{dK bnd} (Q.n)(w.n)
          - int_{dK int} |s| (Q_K - Q_nbr).w               [upwind only]
        The advecting Q* on the facet is taken from the '+' side as in the reference; here we use
        each cell's own trace with its own normal, identical when Q*.n is continuous.
        """
        nc = self.mesh.nc
        gphi = np.einsum("ndc,iqd->niqc", self.Jinv, self.dphiQ)
        Qs_q = self.eval_Q(Qstar)  # [nc,nq,2]
        gradQ = np.einsum("nci,niqd->nqcd", Q, gphi)  # d_d Q_c
        adv = np.einsum("nqd,nqcd->nqc", Qs_q, gradQ)
        out = -self.detJ[:, None, None] * np.einsum("q,nqc,iq->nci", self.wq, adv, self.phiQ)
        Qf = self.eval_Q_facet(Q)
        Qn = self.nbr_facet_values(Qf)
        Qsf = self.eval_Q_facet(Qstar)
        s = self._plus_side_flux(Qsf)  # [nc,3,nqf] = Q*('+').n_K
        interior = (self.nbr >= 0)
        jump = Qf - Qn
        hf = self.hF_inv[self.mesh.cell_facet]  # [nc,3]
        njump = np.einsum("nec,neqc->neq", self.normal, jump)
        nQ_ = np.einsum("nec,neqc->neq", self.normal, Qf)
        coef_vec = np.where(interior[:, :, None], 0.5 * s, 0.0)
        if self.flux == "upwind":
            coef_vec = coef_vec - np.where(interior[:, :, None], np.abs(s), 0.0)
        vec = coef_vec[..., None] * jump
        pen = np.where(interior[:, :, None], njump, nQ_) * (self.alpha * hf)[:, :, None]
        vec = vec - pen[..., None] * self.normal[:, :, None, :]
        out += np.einsum("ne,q,neqc,eiq->nci", self.elen, self.wf, vec, self.phiQ_f)
        return out

    def _plus_side_flux(self, Qsf):
        """Q*('+') . n_K at the facet points of every cell: the '+' side is the first cell of the
        facet (facet_cell[:,0]); the value is expressed with *this* cell's outward normal."""
        own = np.einsum("nec,neqc->neq", self.normal, Qsf)
        nbrv = self.nbr_facet_values(Qsf)
        fc = self.mesh.facet_cell[self.mesh.cell_facet, 0]  # [nc,3] '+' cell of each facet
        is_plus = fc == np.arange(self.mesh.nc)[:, None]
        from_nbr = np.einsum("nec,neqc->neq", self.normal, nbrv)
        return np.where(is_plus[:, :, None], own, from_nbr)

    def f_impl_matrix(self, Qstar):
        """sparse matrix of w,Q -> f_impl(w,Q;Q*) (columns by applying to unit vectors cell-blockwise)"""
        nc, nQ = self.mesh.nc, self.nQ
        n = nc * nQ
        # colouring by local dof: apply to all cells simultaneously for one local dof; contributions
        # to a row block come from the cell itself and its 3 neighbours, which would alias.  Use a
        # distance-2 safe approach instead: probe neighbours separately via masks.
        rows, cols, vals = [], [], []
        colour = _greedy_colouring(self.nbr)
        ncol = colour.max() + 1
        for col in range(ncol):
            mask = colour == col
            cells = np.nonzero(mask)[0]
            for j in range(nQ):
                Q = np.zeros((nc, nQ))
                Q[cells, j] = 1.0
                y = self.f_impl_apply(Q.reshape(nc, 2, self.nQ1), Qstar).reshape(nc, nQ)
                # rows of the active cells themselves
                r = y[cells]
                rr = cells[:, None] * nQ + np.arange(nQ)[None, :]
                rows.append(rr.ravel()); cols.append(np.repeat(cells * nQ + j, nQ)); vals.append(r.ravel())
                for e in range(3):
                    nb = self.nbr[cells, e]
                    ok = nb >= 0
                    r = y[nb[ok]]
                    rr = nb[ok][:, None] * nQ + np.arange(nQ)[None, :]
                    rows.append(rr.ravel()); cols.append(np.repeat(cells[ok] * nQ + j, nQ)); vals.append(r.ravel())
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        return A.tocsr()


def _greedy_colouring(nbr):
    """distance-2 colouring of the cell adjacency graph (cells sharing a neighbour get different colours)"""
    nc = nbr.shape[0]
    colour = np.full(nc, -1, dtype=np.int64)
    for c in range(nc):
        used = set()
        for n1 in nbr[c]:
            if n1 >= 0:
                used.add(colour[n1])
                for n2 in nbr[n1]:
                    if n2 >= 0:
                        used.add(colour[n2])
        col = 0
        while col in used:
            col += 1
        colour[c] = col
    return colour


def _nedelec1_basis(k: int, xq):
    """basis of NED1_k on the reference triangle, [P_{k-1}]^2 (+) S_k with
    S_k = { p in [P~_k]^2 : p . x = 0 } = span{ (-eta, xi) q, q in P~_{k-1} }.   dim = k(k+2).
    Returned as values at xq: [dim, nq, 2].  For k = 0 the space is empty."""
    nq = xq.shape[0]
    if k == 0:
        return np.zeros((0, nq, 2))
    out = []
    if k - 1 >= 0:
        ph = R.dubiner(k - 1, xq)
        for c in range(2):
            for i in range(ph.shape[0]):
                v = np.zeros((nq, 2))
                v[:, c] = ph[i]
                out.append(v)
    xi, eta = xq[:, 0], xq[:, 1]
    for a in range(k):  # homogeneous monomials of degree k-1
        q = xi ** a * eta ** (k - 1 - a)
        out.append(np.stack([-eta * q, xi * q], axis=-1))
    return np.array(out)
