#!/usr/bin/env python
"""Golden vectors of the reference's OWN weak forms, produced by executing the form-building functions of the
reference source through `oracle/miniufl.py` (no Firedrake):

    _f_impl, _pressure_gradient, _Gamma, _weak_divergence   src/timesteppers/hdg_imex.py:313-365
    a_mixed_poisson (action)                                 src/timesteppers/hdg_imex.py:123-127
    _tracer_advection (project_onto_cg=False)                src/timesteppers/common.py:110-129
    a_tentative, b_rhs_tentative, a_poisson, b_rhs_poisson   src/timesteppers/hdg_implicit.py:103-145 (inline in solve)
    a_trace, b_rhs_trace of _reconstruct_trace               src/timesteppers/hdg_imex.py:450-469
    _residual, _final_residual, a_tentative, b_rhs_tentative, b_rhs_mixed_poisson of every IMEX tableau
                                                             src/timesteppers/hdg_imex.py:177-179,233-247,367-413

    python tests/golden/make_golden_forms.py [/root/reference]

The functions are cut out of the reference with `ast` and called with mini-UFL objects on seeded coefficient
data; the assembled dual vectors go to tests/golden/forms_v1.npz together with the inputs.  The tests compare
the hand-written oracle (`oracle/hdg_oracle.py`, `oracle/tracer.py`) with them, so the restatement is checked
against the reference's source text itself; what remains assumed is the meaning of the UFL operators
(documented in oracle/miniufl.py) and the discretisation data.  `/root/reference` is read here only.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from incompressibleeulerhdg_b200.mesh import PeriodicSquareMesh, UnitSquareMesh  # noqa: E402
from oracle.hdg_oracle import HDGOracle  # noqa: E402
from oracle.miniufl import Forms  # noqa: E402

SEED = 123456789


def reference_functions(path, cls_name, names, namespace):
    """compile the named methods of a reference class as plain functions in `namespace`"""
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    out = {}
    for item in cls.body:
        if isinstance(item, ast.FunctionDef) and item.name in names:
            item.decorator_list = []
            mod = ast.fix_missing_locations(ast.Module(body=[item], type_ignores=[]))
            ns = dict(namespace)
            exec(compile(mod, path, "exec"), ns)
            out[item.name] = ns[item.name]
    assert set(out) == set(names), set(names) - set(out)
    return out


def mixed_poisson_expression(path):
    """the right-hand side of `a_mixed_poisson = ...` (`hdg_imex.py:123-127`) as a code object"""
    tree = ast.parse(open(path).read())
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.Assign)
                and getattr(n.targets[0], "id", None) == "a_mixed_poisson")
    return compile(ast.fix_missing_locations(ast.Expression(node.value)), path, "eval")


def case(mesh, k, src, rng, tag, out):
    o = HDGOracle(mesh, k)
    F = Forms(o)
    ns = F.namespace()
    imex = os.path.join(src, "timesteppers", "hdg_imex.py")
    fn = reference_functions(imex, "IncompressibleEulerHDGIMEX",
                             ["_f_impl", "_pressure_gradient", "_Gamma", "_weak_divergence"], ns)
    tr = reference_functions(os.path.join(src, "timesteppers", "common.py"), "IncompressibleEuler",
                             ["_tracer_advection"], ns)
    nc, nf = mesh.nc, mesh.nf
    Q = rng.standard_normal((nc, 2, o.nQ1))
    Qs_raw = rng.standard_normal((nc, 2, o.nQ1))
    Qs = o.project_bdm(Qs_raw)
    p = rng.standard_normal((nc, o.np_))
    lam = rng.standard_normal((nf, o.nl1))
    q = rng.standard_normal((nc, o.np_))
    w, psi, mu, chi = F.test("Q"), F.test("P"), F.test("T"), F.test("P")
    cQ, cQs, cQsr, cp, cl, cq = (F.coefficient(s, d) for s, d in
                                 (("Q", Q), ("Q", Qs), ("Q", Qs_raw), ("P", p), ("T", lam), ("P", q)))
    put = lambda name, arr: out.__setitem__(f"{tag}/{name}", arr)
    for name, arr in (("Q", Q), ("Qstar_raw", Qs_raw), ("p", p), ("lam", lam), ("q", q)):
        put("in_" + name, arr)
    for flux in ("upwind", "centered"):
        self = types.SimpleNamespace(_mesh=None, alpha_penalty=1, _hF_inv=F.hF_inv(), flux=flux, tau=1.0)
        put(f"f_impl_{flux}", F.assemble(fn["_f_impl"](self, w, cQ, cQs))["Q"])
        put(f"f_impl_{flux}_rawQstar", F.assemble(fn["_f_impl"](self, w, cQ, cQsr))["Q"])
    self = types.SimpleNamespace(_mesh=None, alpha_penalty=1, _hF_inv=F.hF_inv(), flux="upwind", tau=1.0,
                                 _pressure_gradient=None, _Gamma=None)
    self._pressure_gradient = lambda *a: fn["_pressure_gradient"](self, *a)
    self._Gamma = lambda *a: fn["_Gamma"](self, *a)
    put("pressure_gradient", F.assemble(fn["_pressure_gradient"](self, w, cp, cl))["Q"])
    g = F.assemble(fn["_Gamma"](self, psi, mu, cQ, cp, cl))
    put("Gamma_P", g["P"])
    put("Gamma_T", g["T"])
    put("weak_divergence", F.assemble(fn["_weak_divergence"](self, psi, cQ))["P"])
    # the mixed-Poisson operator of hdg_imex.py:123-127 applied to (u, phi, lmbda) = (Q, p, lam)
    env = dict(ns, self=self, w=w, psi=psi, mu=mu, u=cQ, phi=cp, lmbda=cl)
    a = F.assemble(eval(mixed_poisson_expression(imex), env))
    put("mixed_poisson_Q", a["Q"])
    put("mixed_poisson_P", a["P"])
    put("mixed_poisson_T", a["T"])
    # tracer: the reference projects the velocity onto CG first (project_onto_cg=True needs a Firedrake solve), so the
    # form is executed with project_onto_cg=False on a *continuous* velocity -- the nodal interpolant of a smooth
    # (2 pi periodic) field, which is what the form sees after the projection
    U = o.interpolate_cell(lambda x, y: (np.sin(x) * np.cos(2 * y) + 0.3, np.cos(x) - 0.5 * np.sin(y) * np.cos(x)), "Q")
    put("in_Ucont", U)
    put("tracer_advection", F.assemble(tr["_tracer_advection"](self, chi, cq, F.coefficient("Q", U),
                                                               project_onto_cg=False))["P"])


def trace_case(mesh, k, src, rng, tag, out):
    """`_reconstruct_trace` (`hdg_imex.py:450-469`): the right-hand side b_rhs_trace and the action of a_trace"""
    o = HDGOracle(mesh, k)
    F = Forms(o)
    path = os.path.join(src, "timesteppers", "hdg_imex.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "IncompressibleEulerHDGIMEX")
    meth = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_reconstruct_trace")

    def expr_of(name):
        node = next(n for n in ast.walk(meth) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == name)
        return compile(ast.fix_missing_locations(ast.Expression(node.value)), path, "eval")

    Q = rng.standard_normal((mesh.nc, 2, o.nQ1))
    p = rng.standard_normal((mesh.nc, o.np_))
    lam = rng.standard_normal((mesh.nf, o.nl1))
    ns = F.namespace()
    env = dict(ns, self=types.SimpleNamespace(tau=1.0), n=ns["FacetNormal"](None), Q=F.coefficient("Q", Q),
               p=F.coefficient("P", p), lmbda=F.coefficient("T", lam), mu=F.test("T"))
    out[f"{tag}/in_Q"], out[f"{tag}/in_p"], out[f"{tag}/in_lam"] = Q, p, lam
    out[f"{tag}/b_rhs_trace"] = F.assemble(eval(expr_of("b_rhs_trace"), env))["T"]
    out[f"{tag}/a_trace_action"] = F.assemble(eval(expr_of("a_trace"), env))["T"]


def chorin_case(mesh, k, src, rng, tag, out):
    """the forms written inline in `IncompressibleEulerHDGImplicit.solve` (`hdg_implicit.py:103-145`):
    a_tentative (with its upwind `+=`), b_rhs_tentative, a_poisson, b_rhs_poisson -- as actions"""
    o = HDGOracle(mesh, k)
    F = Forms(o)
    path = os.path.join(src, "timesteppers", "hdg_implicit.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "IncompressibleEulerHDGImplicit")
    solve = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "solve")

    def expr_of(name, aug=False):
        kind = ast.AugAssign if aug else ast.Assign
        for n in ast.walk(solve):
            if isinstance(n, kind):
                tgt = n.target if aug else n.targets[0]
                if getattr(tgt, "id", None) == name:
                    return compile(ast.fix_missing_locations(ast.Expression(n.value)), path, "eval")
        raise KeyError(name)

    nc, nf, dt = mesh.nc, mesh.nf, 0.02
    X, Q, f = (rng.standard_normal((nc, 2, o.nQ1)) for _ in range(3))
    Qs = o.project_bdm(rng.standard_normal((nc, 2, o.nQ1)))
    p, lam = rng.standard_normal((nc, o.np_)), rng.standard_normal((nf, o.nl1))
    put = lambda name, arr: out.__setitem__(f"{tag}/{name}", arr)
    for name, arr in (("X", X), ("Q", Q), ("f", f), ("Qstar", Qs), ("p", p), ("lam", lam)):
        put("in_" + name, arr)
    put("in_dt", np.array(dt))
    self = types.SimpleNamespace(_dt=dt, alpha=1, _hF_inv=F.hF_inv(), tau=1.0, flux="upwind")
    env = dict(F.namespace(), self=self, n=F.namespace()["FacetNormal"](None), u_Q=F.coefficient("Q", X), w_Q=F.test("Q"),
               Q_star=F.coefficient("Q", Qs), Q=F.coefficient("Q", Q), f=F.coefficient("Q", f),
               u=F.coefficient("Q", X), phi=F.coefficient("P", p), lmbda=F.coefficient("T", lam),
               w=F.test("Q"), psi=F.test("P"), mu=F.test("T"), Q_tentative=F.coefficient("Q", X))
    a_cen = eval(expr_of("a_tentative"), env)
    put("a_tentative_centered", F.assemble(a_cen)["Q"])
    put("a_tentative_upwind", F.assemble(a_cen + eval(expr_of("a_tentative", aug=True), env))["Q"])
    put("b_rhs_tentative", F.assemble(eval(expr_of("b_rhs_tentative"), env))["Q"])
    a = F.assemble(eval(expr_of("a_poisson"), env))
    put("a_poisson_Q", a["Q"])
    put("a_poisson_P", a["P"])
    put("a_poisson_T", a["T"])
    put("b_rhs_poisson", F.assemble(eval(expr_of("b_rhs_poisson"), env))["P"])


def imex_case(mesh, k, src, rng, cls_name, tab, tag, out):
    """the IMEX stage forms: `_residual(w, i)` / `_final_residual(w)` (`hdg_imex.py:367-413`, methods) and, from the
    loops of `__init__`, `a_tentative` (as an action), `b_rhs_tentative` (`:233-247`) and `b_rhs_mixed_poisson`
    (`:177-179`), for the tableau data of `cls_name` (tests/golden/tableaux_v1.json, itself executed from the reference)"""
    o = HDGOracle(mesh, k)
    F = Forms(o)
    ns = F.namespace()
    imex = os.path.join(src, "timesteppers", "hdg_imex.py")
    names = ["_f_impl", "_pressure_gradient", "_weak_divergence", "_residual", "_final_residual"]
    ns["split"] = lambda state: state
    fn = reference_functions(imex, "IncompressibleEulerHDGIMEX", names, ns)
    tree = ast.parse(open(imex).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "IncompressibleEulerHDGIMEX")
    init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")

    def expr_of(name):
        node = next(n for n in ast.walk(init) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == name
                    and not isinstance(n.value, ast.Dict))
        return compile(ast.fix_missing_locations(ast.Expression(node.value)), imex, "eval")

    s_, nc, nf, dt = tab["nstages"], mesh.nc, mesh.nf, 0.02
    rQ = lambda: rng.standard_normal((nc, 2, o.nQ1))
    stage = [(rQ(), rng.standard_normal((nc, o.np_)), rng.standard_normal((nf, o.nl1))) for _ in range(s_)]
    b_rhs = [rQ() for _ in range(s_)]
    Qstar = [o.project_bdm(rQ()) for _ in range(s_ - 1)]
    Qtent = [rQ() for _ in range(s_)]
    X = rQ()
    put = lambda name, arr: out.__setitem__(f"{tag}/{name}", np.asarray(arr))
    put("in_dt", dt)
    put("in_X", X)
    for j in range(s_):
        put(f"in_stage{j}_Q", stage[j][0]); put(f"in_stage{j}_p", stage[j][1]); put(f"in_stage{j}_l", stage[j][2])
        put(f"in_b_rhs{j}", b_rhs[j]); put(f"in_Qtent{j}", Qtent[j])
        if j < s_ - 1:
            put(f"in_Qstar{j}", Qstar[j])
    self = types.SimpleNamespace(
        _mesh=None, alpha_penalty=1, _hF_inv=F.hF_inv(), flux="upwind", tau=1.0, _dt=dt, nstages=s_,
        _a_expl=np.asarray(tab["_a_expl"]), _a_impl=np.asarray(tab["_a_impl"]), _b_expl=np.asarray(tab["_b_expl"]),
        _b_impl=np.asarray(tab["_b_impl"]), _c_expl=np.asarray(tab["_c_expl"]),
        _stage_state=[tuple(F.coefficient(sp_, d) for sp_, d in zip(("Q", "P", "T"), st)) for st in stage],
        _b_rhs=[F.coefficient("Q", b) for b in b_rhs], _Qstar=[F.coefficient("Q", q) for q in Qstar],
        _Q_tentative=[F.coefficient("Q", q) for q in Qtent])
    for name in names:
        setattr(self, name, (lambda f: (lambda *a, **kw: f(self, *a, **kw)))(fn[name]))
    w, psi = F.test("Q"), F.test("P")
    put("final_residual", F.assemble(self._final_residual(w))["Q"])
    for i in range(1, s_):
        put(f"residual_{i}", F.assemble(self._residual(w, i))["Q"])
        Q_i, p_i, lambda_i = self._stage_state[i]
        env = dict(ns, self=self, i=i, w_Q=w, u_Q=F.coefficient("Q", X), Q_i=Q_i, p_i=p_i, lambda_i=lambda_i, psi=psi)
        put(f"a_tentative_{i}", F.assemble(eval(expr_of("a_tentative"), env))["Q"])
        put(f"b_rhs_tentative_{i}", F.assemble(eval(expr_of("b_rhs_tentative"), env))["Q"])
        put(f"b_rhs_mixed_poisson_{i}", F.assemble(eval(expr_of("b_rhs_mixed_poisson"), env))["P"])


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(root, "src")
    out = {}
    rng = np.random.default_rng(SEED)
    case(UnitSquareMesh(3, perturb=0.15), 1, src, rng, "square_k1", out)
    case(UnitSquareMesh(3, perturb=0.15), 2, src, rng, "square_k2", out)
    case(PeriodicSquareMesh(3, L=2 * np.pi), 2, src, rng, "periodic_k2", out)
    chorin_case(UnitSquareMesh(3, perturb=0.15), 2, src, rng, "chorin_square_k2", out)
    chorin_case(PeriodicSquareMesh(3, L=2 * np.pi), 1, src, rng, "chorin_periodic_k1", out)
    trace_case(UnitSquareMesh(3, perturb=0.15), 2, src, rng, "trace_square_k2", out)
    trace_case(PeriodicSquareMesh(3, L=2 * np.pi), 1, src, rng, "trace_periodic_k1", out)
    import json

    tabs = json.load(open(os.path.join(HERE, "tableaux_v1.json")))["classes"]
    for cls_name, tab in sorted(tabs.items()):
        imex_case(UnitSquareMesh(3, perturb=0.15), 1, src, rng, cls_name, tab, "imex_" + cls_name, out)
    path = os.path.join(HERE, "forms_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")
