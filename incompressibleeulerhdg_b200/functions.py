"""Minimal stand-ins for the Firedrake objects the reference timesteppers touch: expressions,
function spaces and functions.  Fields live on the GPU (torch tensors in the engine's SoA layout);
nothing here does arithmetic on the hot path -- it only marshals data and calls the C-ABI.

* ``Expression``     a callable f(x, y) (scalar or 2-vector) times a scalar factor.  Scalar
                     multiples share the interpolated base field, so time-dependent forcings
                     ``Psi'(t) * Q_s`` (`model_problems.py:71-80`) are interpolated once.
* ``FunctionSpace``  DG_{k+1}^2 ("Q"), DG_k ("p") or DGT_k ("trace") on a mesh, bound to an engine.
* ``Function``       coefficient tensor + space.

Interpolation follows Firedrake's ``Function.interpolate`` (`hdg_imex.py:520-521`): point
evaluation at the Lagrange nodes of the space (equispaced by default -- the node set is a
parameter because Firedrake's default variant differs between versions), then
conversion to the engine's modal basis with the reference-element Vandermonde matrix.
"""

from __future__ import annotations

import numpy as np

from . import refelem as R

__all__ = ["Expression", "FunctionSpace", "Function", "as_expression"]


class Expression:
    def __init__(self, fun, rank: int, scale: float = 1.0, base=None):
        self.fun = fun
        self.rank = rank  # 0 scalar, 1 vector
        self.scale = float(scale)
        self.base = self if base is None else base

    def __mul__(self, c):
        return Expression(self.fun, self.rank, self.scale * float(c), base=self.base)

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    def __call__(self, x, y):
        v = self.fun(x, y)
        if self.rank == 1:
            return tuple(self.scale * np.broadcast_to(c, np.shape(x)) for c in v)
        return self.scale * np.broadcast_to(v, np.shape(x))


def as_expression(obj, rank):
    if isinstance(obj, Expression):
        return obj
    if callable(obj):
        return Expression(obj, rank)
    if rank == 0:
        c = float(obj)
        return Expression(lambda x, y: c + 0 * x, 0)
    cx, cy = (float(v) for v in obj)
    return Expression(lambda x, y: (cx + 0 * x, cy + 0 * x), 1)


class FunctionSpace:
    KIND = {"Q": 0, "p": 1, "trace": 2}

    def __init__(self, engine, name: str, nodes=None):
        self.engine = engine
        self.name = name
        self.kind = self.KIND[name]
        k = engine.k
        self.degree = k + 1 if name == "Q" else k
        self._cache = {}
        if name in ("Q", "p"):
            self.nodes = R.lagrange_nodes_cell(self.degree) if nodes is None else np.asarray(nodes)
            self.Vinv = R.nodal_to_modal_cell(self.degree, self.nodes)
        else:
            self.nodes = R.lagrange_nodes_facet(self.degree) if nodes is None else np.asarray(nodes)
            self.Vinv = R.nodal_to_modal_facet(self.degree, self.nodes)

    def mesh(self):
        return self.engine.mesh

    # physical coordinates of the nodes of every cell [nc, nnodes, 2]
    def node_coordinates(self):
        assert self.name in ("Q", "p")
        xy = self.engine.mesh.cell_xy
        x0 = xy[:, 0]
        J = np.stack([xy[:, 1] - xy[:, 0], xy[:, 2] - xy[:, 0]], axis=-1)  # [nc, c, d]
        return x0[:, None, :] + np.einsum("ncd,qd->nqc", J, self.nodes)

    def _interpolate_base(self, expr: Expression):
        key = id(expr.base)
        hit = self._cache.get(key)
        if hit is not None and hit[0] is expr.base:
            return hit[1]
        xp = self.node_coordinates()
        base = expr.base
        if self.name == "Q":
            v = base.fun(xp[..., 0], xp[..., 1])
            v = np.stack([np.broadcast_to(c, xp.shape[:2]) for c in v], axis=-1)  # [nc, q, 2]
            coef = np.einsum("iq,nqc->nci", self.Vinv, v)
        else:
            v = np.broadcast_to(base.fun(xp[..., 0], xp[..., 1]), xp.shape[:2])
            coef = np.einsum("aq,nq->na", self.Vinv, v)
        dev = self.engine.upload(self.kind, coef)
        self._cache[key] = (base, dev)
        return dev

    def interpolate(self, expr, out=None):
        """Function(space).interpolate(expr): nodal interpolation, evaluated on the host once per
        base expression and scaled on the device"""
        expr = as_expression(expr, 1 if self.name == "Q" else 0)
        base = self._interpolate_base(expr)
        f = Function(self) if out is None else out
        self.engine.lincomb_dev(f.data, [(expr.scale, base)])
        return f

    def zeros(self):
        f = Function(self)
        f.data.zero_()
        return f


class Function:
    def __init__(self, space: FunctionSpace, data=None, name: str | None = None):
        self.space = space
        self.data = space.engine.empty(space.kind) if data is None else data
        self.name = name

    def function_space(self):
        return self.space

    def rename(self, name):
        self.name = name

    def assign(self, other):
        self.data.copy_(other.data)
        return self

    def copy(self):
        f = Function(self.space, name=self.name)
        f.data.copy_(self.data)
        return f

    def to_host(self):
        """coefficients in the host AoS layout (modal basis)"""
        return self.space.engine.download(self.space.kind, self.data)
