"""A minimal UFL interpreter  --  TEST INFRASTRUCTURE ONLY (same rules as hdg_oracle.py).

Purpose: execute the reference's OWN form-building functions (`_f_impl`, `_pressure_gradient`, `_Gamma`,
`_weak_divergence` of `src/timesteppers/hdg_imex.py:313-365`, `_tracer_advection` of
`src/timesteppers/common.py:110-129`) without Firedrake, so that the oracle's hand-written restatement of
those forms is checked against the reference's source text itself rather than against a second reading of
it.  `tests/golden/make_golden_forms.py` cuts the functions out of the reference with `ast`, runs them on
the objects defined here with seeded coefficient data, and stores the assembled vectors as golden files.

What is taken from the reference: every sign, factor, restriction and measure of the forms.  What is supplied
here (documented UFL semantics, [FD-knowledge]): the meaning of the operators --

  dx / ds / dS          cell, exterior-facet, interior-facet integrals
  X('+'), X('-')        restriction to the two cells of an interior facet; FacetNormal('-') = -FacetNormal('+')
  avg(X)                (X('+') + X('-')) / 2
  grad(v)[i, j]         d v_i / d x_j ;  div(v) = sum_i d v_i / d x_i ;  div(s v) = grad(s) . v + s div(v)
  inner / outer         full contraction / dyadic product ;  abs() pointwise

-- and the discretisation data of `HDGOracle` (bases, quadrature, geometry).  Only *actions* are assembled
(forms that are linear in one test function, with every other argument a coefficient), which is all that
pinning the restatement needs.

Value arrays have the layout [entity, point, *tensor, test] with a broadcastable test axis (length 1 for
test-free expressions).  On dS the test axis covers the dofs of both cells ('+' first) for cell spaces and
the facet's own dofs for the trace space.
"""

from __future__ import annotations

import numpy as np

__all__ = ["Forms"]


class Expr:
    rank = 0  # tensor rank

    # -- arithmetic ---------------------------------------------------------------------------
    def __add__(self, o):
        return Sum(self, as_expr(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Sum(self, Scale(as_expr(o), -1.0))

    def __rsub__(self, o):
        return Sum(as_expr(o), Scale(self, -1.0))

    def __neg__(self):
        return Scale(self, -1.0)

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([(self, o)])
        if isinstance(o, Form):  # Constant(c) * (form + form)
            return o * self
        return Product(self, as_expr(o))

    def __rmul__(self, o):
        return Product(as_expr(o), self)

    def __truediv__(self, o):
        return Scale(self, 1.0 / float(o))

    def __abs__(self):
        return Abs(self)

    def __call__(self, side):
        assert side in ("+", "-")
        return Restricted(self, side)


class Const(Expr):
    def __init__(self, v):
        self.v = float(v)

    def ev(self, ctx, side):
        return np.full((1, 1, 1), self.v), None

    def __float__(self):
        return self.v


def as_expr(o):
    return o if isinstance(o, Expr) else Const(o)


def _merge_test(ta, tb):
    assert ta is None or tb is None, "a form must be linear in the test function"
    return ta if ta is not None else tb


class Sum(Expr):
    def __init__(self, a, b):
        assert a.rank == b.rank or isinstance(a, Const) or isinstance(b, Const), "rank mismatch in a sum"
        self.a, self.b, self.rank = a, b, max(a.rank, b.rank)

    def ev(self, ctx, side):
        (va, ta), (vb, tb) = self.a.ev(ctx, side), self.b.ev(ctx, side)
        if ta is not None and tb is not None:
            assert ta == tb, "sum of different test spaces inside one integrand"
            return va + vb, ta
        # a test-free term cannot be added to a test term inside an integrand (not linear)
        assert ta is None and tb is None, "sum of a test and a test-free expression"
        return va + vb, None


class Scale(Expr):
    def __init__(self, a, c):
        self.a, self.c, self.rank = a, float(c), a.rank

    def ev(self, ctx, side):
        v, t = self.a.ev(ctx, side)
        return self.c * v, t


class Product(Expr):
    """at least one operand is scalar valued"""

    def __init__(self, a, b):
        assert a.rank == 0 or b.rank == 0, "only scalar * tensor products occur in the reference forms"
        self.a, self.b, self.rank = a, b, max(a.rank, b.rank)

    def ev(self, ctx, side):
        (va, ta), (vb, tb) = self.a.ev(ctx, side), self.b.ev(ctx, side)
        r = self.rank
        # insert singleton tensor axes into the scalar operand: [..., test] -> [..., 1*r, test]
        if self.a.rank < r:
            va = va.reshape(va.shape[:2] + (1,) * r + va.shape[2:])
        if self.b.rank < r:
            vb = vb.reshape(vb.shape[:2] + (1,) * r + vb.shape[2:])
        return va * vb, _merge_test(ta, tb)


class Abs(Expr):
    def __init__(self, a):
        assert a.rank == 0
        self.a = a

    def ev(self, ctx, side):
        v, t = self.a.ev(ctx, side)
        assert t is None, "abs of a test function"
        return np.abs(v), None


class Inner(Expr):
    def __init__(self, a, b):
        assert a.rank == b.rank
        self.a, self.b = a, b

    def ev(self, ctx, side):
        (va, ta), (vb, tb) = self.a.ev(ctx, side), self.b.ev(ctx, side)
        r = self.a.rank
        axes = tuple(range(2, 2 + r))
        return (va * vb).sum(axis=axes) if r else va * vb, _merge_test(ta, tb)


class Outer(Expr):
    rank = 2

    def __init__(self, a, b):
        assert a.rank == 1 and b.rank == 1
        self.a, self.b = a, b

    def ev(self, ctx, side):
        (va, ta), (vb, tb) = self.a.ev(ctx, side), self.b.ev(ctx, side)
        return va[:, :, :, None, :] * vb[:, :, None, :, :], _merge_test(ta, tb)


class Restricted(Expr):
    def __init__(self, a, side):
        self.a, self.side, self.rank = a, side, a.rank

    def ev(self, ctx, side):
        assert ctx.kind == "int", "restrictions only make sense under dS"
        return self.a.ev(ctx, self.side)


class Avg(Expr):
    def __init__(self, a):
        self.a, self.rank = a, a.rank

    def ev(self, ctx, side):
        assert ctx.kind == "int"
        (vp, tp), (vm, tm) = self.a.ev(ctx, "+"), self.a.ev(ctx, "-")
        assert tp == tm
        return 0.5 * (vp + vm), tp


class Grad(Expr):
    def __init__(self, a):
        assert isinstance(a, (Coefficient, TestFunction)), "grad of a terminal only"
        self.a, self.rank = a, a.rank + 1

    def ev(self, ctx, side):
        return self.a.ev_grad(ctx, side)


class Div(Expr):
    rank = 0

    def __init__(self, a):
        assert a.rank == 1
        self.a = a

    def ev(self, ctx, side):
        a = self.a
        if isinstance(a, Product):  # div(s v) = grad(s) . v + s div(v)
            s, v = (a.a, a.b) if a.a.rank == 0 else (a.b, a.a)
            return Sum(Inner(Grad(s), v), Product(s, Div(v))).ev(ctx, side)
        g, t = a.ev_grad(ctx, side)  # [ent, q, i, j, test]
        return g[:, :, 0, 0] + g[:, :, 1, 1], t


class FacetNormalExpr(Expr):
    rank = 1

    def __init__(self, forms):
        self.f = forms

    def ev(self, ctx, side):
        return self.f._normal(ctx, side)[..., None], None


class Coefficient(Expr):
    def __init__(self, forms, space, data):
        self.f, self.space, self.data = forms, space, np.asarray(data, dtype=float)
        self.rank = 1 if space == "Q" else 0

    def ev(self, ctx, side):
        return self.f._coef(self, ctx, side, grad=False)[..., None], None

    def ev_grad(self, ctx, side):
        return self.f._coef(self, ctx, side, grad=True)[..., None], None


class TestFunction(Expr):
    def __init__(self, forms, space):
        self.f, self.space = forms, space
        self.rank = 1 if space == "Q" else 0

    def ev(self, ctx, side):
        return self.f._test(self, ctx, side, grad=False), self.space

    def ev_grad(self, ctx, side):
        return self.f._test(self, ctx, side, grad=True), self.space


class Measure:
    def __init__(self, kind):
        self.kind = kind


class Form:
    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        return Form(self.integrals + o.integrals)

    def __sub__(self, o):
        return Form(self.integrals + [(Scale(e, -1.0), m) for e, m in o.integrals])

    def __neg__(self):
        return Form([(Scale(e, -1.0), m) for e, m in self.integrals])

    def __rmul__(self, c):
        return Form([(Scale(e, float(c)), m) for e, m in self.integrals])

    __mul__ = __rmul__


class _Ctx:
    def __init__(self, kind):
        self.kind = kind


class Forms:
    """the UFL names the reference's form functions use, bound to one `HDGOracle`"""

    def __init__(self, oracle):
        self.o = o = oracle
        m = o.mesh
        self.dx, self.ds, self.dS = Measure("cell"), Measure("ext"), Measure("int")
        fc, fl = m.facet_cell, m.facet_local
        self.int_f = np.nonzero(fc[:, 1] >= 0)[0]
        self.ext_f = np.nonzero(fc[:, 1] < 0)[0]
        # (cell, local facet) of the sides: '+' = first cell of the facet (the engine's and the oracle's convention;
        # the reference's forms are symmetric under swapping the sides)
        self.side = {"+": (fc[self.int_f, 0], fl[self.int_f, 0]), "-": (fc[self.int_f, 1], fl[self.int_f, 1]),
                     None: (fc[self.ext_f, 0], fl[self.ext_f, 0])}

    # -- public constructors (the names a form function sees) -----------------------------------------
    def namespace(self):
        return dict(inner=Inner, outer=Outer, grad=Grad, div=Div, avg=Avg, dx=self.dx, ds=self.ds, dS=self.dS,
                    FacetNormal=lambda mesh: FacetNormalExpr(self), Constant=lambda v: Const(v), abs=abs)

    def coefficient(self, space, data):
        return Coefficient(self, space, data)

    def test(self, space):
        return TestFunction(self, space)

    def hF_inv(self):
        """the DGT0 field 1/h_F of `common.py:36-57`"""
        return Coefficient(self, "F0", self.o.hF_inv)

    # -- tabulation helpers ----------------------------------------------------------------------------------
    def _cells_facets(self, ctx, side):
        if ctx.kind == "int":
            assert side in ("+", "-"), "unrestricted cell-wise quantity under dS"
            return self.side[side]
        return self.side[None]

    def _facet_tab(self, tab, e, reverse):
        """tab [3, ndof, nqf(, 2)] -> [nent, ndof, nqf(, 2)] at the points of the '+' side"""
        t = tab[e]
        return t[:, :, ::-1] if reverse else t

    def _normal(self, ctx, side):
        o = self.o
        if ctx.kind == "cell":
            raise AssertionError("FacetNormal under dx")
        c, e = self._cells_facets(ctx, side)
        return o.normal[c, e][:, None, :]  # [ent, 1, 2]; the '-' cell's outward normal is the opposite vector

    def _basis(self, space, ctx, side):
        """(values [ent, ndof, q], cells, dof data lookup) of the cell-wise basis at the integration points"""
        o = self.o
        phi, phif = (o.phiQ, o.phiQ_f) if space == "Q" else (o.phiP, o.phiP_f)
        if ctx.kind == "cell":
            return np.broadcast_to(phi[None], (o.mesh.nc,) + phi.shape), np.arange(o.mesh.nc)
        c, e = self._cells_facets(ctx, side)
        return self._facet_tab(phif, e, reverse=(side == "-")), c

    def _trace_basis(self, ctx, side):
        """Legendre basis along the global facet direction at the integration points [ent, k+1, q]"""
        o = self.o
        assert ctx.kind != "cell", "trace function under dx"
        c, e = self.side["+"] if ctx.kind == "int" else self.side[None]  # single valued: evaluate from the '+' cell
        return o.ell[o.mesh.cell_flip[c, e]], (self.int_f if ctx.kind == "int" else self.ext_f)

    def _phys_grad(self, space, cells):
        o = self.o
        d = o.dphiQ if space == "Q" else o.dphiP  # [ndof, q, 2]
        return np.einsum("ndc,iqd->niqc", o.Jinv[cells], d)  # [ent, ndof, q, c]

    # -- terminals -------------------------------------------------------------------------------------------------
    def _coef(self, u, ctx, side, grad):
        o = self.o
        if u.space == "F0":
            assert not grad
            f = self.int_f if ctx.kind == "int" else self.ext_f
            return u.data[f][:, None]
        if u.space == "T":
            assert not grad
            b, f = self._trace_basis(ctx, side)
            return np.einsum("fm,fmq->fq", u.data[f], b)
        if grad:
            assert ctx.kind == "cell", "derivatives of coefficients are only needed under dx"
            g = self._phys_grad(u.space, np.arange(o.mesh.nc))
            if u.space == "Q":
                return np.einsum("nci,niqd->nqcd", u.data, g)  # grad(Q)[c, d] = d_d Q_c
            return np.einsum("na,naqd->nqd", u.data, g)
        b, cells = self._basis(u.space, ctx, side)
        if u.space == "Q":
            return np.einsum("nci,niq->nqc", u.data[cells], b)
        return np.einsum("na,naq->nq", u.data[cells], b)

    def _test(self, w, ctx, side, grad):
        """[ent, q, *tensor, ntest]"""
        o = self.o
        if w.space == "T":
            assert not grad
            b, _ = self._trace_basis(ctx, side)
            return np.swapaxes(b, 1, 2)  # [ent, q, k+1]
        nloc = o.nQ1 if w.space == "Q" else o.np_
        if grad:
            assert ctx.kind == "cell"
            g = self._phys_grad(w.space, np.arange(o.mesh.nc))  # [n, a, q, d]
            if w.space == "P":
                return np.einsum("naqd->nqda", g)
            out = np.zeros((o.mesh.nc, g.shape[2], 2, 2, 2 * nloc))  # grad(w)[c, d] for test dof (c', a)
            for c in range(2):
                out[:, :, c, :, c * nloc:(c + 1) * nloc] = np.einsum("naqd->nqda", g)
            return out
        b, _ = self._basis(w.space, ctx, side)  # [ent, a, q]
        vals = np.swapaxes(b, 1, 2)  # [ent, q, a]
        if w.space == "Q":
            v = np.zeros(vals.shape[:2] + (2, 2 * nloc))
            for c in range(2):
                v[:, :, c, c * nloc:(c + 1) * nloc] = vals
            vals = v
        if ctx.kind == "int":  # test axis = dofs of the '+' cell followed by those of the '-' cell
            n = vals.shape[-1]
            both = np.zeros(vals.shape[:-1] + (2 * n,))
            both[..., (0 if side == "+" else n):(n if side == "+" else 2 * n)] = vals
            vals = both
        return vals

    # -- assembly --------------------------------------------------------------------------------------------------
    def assemble(self, form):
        """dual vectors of a linear form: dict space -> array (Q [nc,2,nQ1], P [nc,np], T [nf,k+1])"""
        o = self.o
        m = o.mesh
        out = {}
        for expr, meas in form.integrals:
            ctx = _Ctx(meas.kind)
            assert expr.rank == 0, "integrand must be scalar"
            v, space = expr.ev(ctx, None)
            assert space is not None, "integrand without a test function"
            if meas.kind == "cell":
                v = np.broadcast_to(v, (m.nc, o.wq.size, v.shape[2]))
                loc = np.einsum("q,n,nqt->nt", o.wq, o.detJ, v)
                ents = [np.arange(m.nc)]
            else:
                c, e = self.side["+"] if meas.kind == "int" else self.side[None]
                v = np.broadcast_to(v, (c.size, o.wf.size, v.shape[2]))
                loc = np.einsum("q,f,fqt->ft", o.wf, o.elen[c, e], v)
                ents = [self.side["+"][0], self.side["-"][0]] if meas.kind == "int" else [self.side[None][0]]
            if space == "T":
                f = self.int_f if meas.kind == "int" else self.ext_f
                acc = out.setdefault("T", np.zeros((m.nf, o.nl1)))
                np.add.at(acc, f, loc)
                continue
            nloc = 2 * o.nQ1 if space == "Q" else o.np_
            acc = out.setdefault(space, np.zeros((m.nc, nloc)))
            for j, cells in enumerate(ents):
                np.add.at(acc, cells, loc[:, j * nloc:(j + 1) * nloc])
        if "Q" in out:
            out["Q"] = out["Q"].reshape(m.nc, 2, o.nQ1)
        return out
