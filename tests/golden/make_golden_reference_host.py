#!/usr/bin/env python
"""Golden outputs of the host-side pieces of the reference that CAN be executed here (no Firedrake needed), so
that the mirrors in incompressibleeulerhdg_b200/ are pinned against the reference itself:

  log_summary      `src/auxilliary/logging.py:34-60` imported as is; the printed table for a fixed set of timings
  averager         class `Averager` of `src/auxilliary/utils.py:11-46` (cut out with ast: the module imports firedrake)
  shear_fourier    the 28 `integrate.quad(...)` Fourier coefficients of the double-shear-layer pressure,
                   `src/model_problems.py:166-179` (the call expression is cut out with ast and evaluated)

    python tests/golden/make_golden_reference_host.py [/root/reference]

Output: tests/golden/reference_host_v1.json.  `/root/reference` is read here only, never by the tests.
"""
import ast
import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TIMINGS = {"timestep": [0.5, 0.25, 0.125], "pressure_solve": [0.03125, 0.0625], "tentative_velocity_solve": [0.75],
           "bdm_projection": [0.001953125, 0.00390625, 0.0009765625, 0.0078125]}
SAMPLES = [3, 7, 7, 12, 5.5, 0, 41]


def golden_log_summary(src):
    spec = importlib.util.spec_from_file_location("ref_logging", os.path.join(src, "auxilliary", "logging.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.PerformanceLog.data.clear()
    for k, v in TIMINGS.items():
        mod.PerformanceLog.data[k].extend(v)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mod.log_summary()
    return buf.getvalue()


def golden_averager(src):
    tree = ast.parse(open(os.path.join(src, "auxilliary", "utils.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Averager")
    ns = {"np": np}
    exec(compile(ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[])), "utils.py", "exec"), ns)
    a = ns["Averager"]()
    out = []
    for x in SAMPLES:
        a.update(x)
        out.append([a.n_samples, float(a.value)])
    text = repr(a)
    a.reset()
    return {"running": out, "repr": text, "after_reset": [a.n_samples, float(a.value)]}


def golden_shear(src):
    import scipy.integrate as integrate

    tree = ast.parse(open(os.path.join(src, "model_problems.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DoubleLayerShearFlow")
    init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
    quad = next(n for n in ast.walk(init) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
                and n.func.attr == "quad")
    kmax = next(n.value.value for n in ast.walk(init) if isinstance(n, ast.Assign)
                and getattr(n.targets[0], "id", None) == "kmax")
    defaults = {a.arg: d for a, d in zip(init.args.args[-len(init.args.defaults):], init.args.defaults)}
    rho = eval(compile(ast.Expression(defaults["rho"]), "model_problems.py", "eval"), {"np": np})
    delta = eval(compile(ast.Expression(defaults["delta"]), "model_problems.py", "eval"), {"np": np})
    code = compile(ast.fix_missing_locations(ast.Expression(quad)), "model_problems.py", "eval")
    coefs = [float(eval(code, {"np": np, "integrate": integrate, "self": types.SimpleNamespace(rho=rho), "k": k})[0])
             for k in range(kmax)]
    return {"kmax": kmax, "rho": float(rho), "delta": float(delta), "fourier_coefficient": coefs}


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(root, "src")
    data = {"timings": TIMINGS, "log_summary": golden_log_summary(src), "samples": SAMPLES,
            "averager": golden_averager(src), "shear": golden_shear(src)}
    with open(os.path.join(HERE, "reference_host_v1.json"), "w") as fh:
        json.dump(data, fh, indent=1)
    print(data["log_summary"])
    print(data["averager"]["repr"], data["shear"]["fourier_coefficient"][:3])
