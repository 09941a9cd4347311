#!/usr/bin/env python
"""Golden IMEX tableaux, taken from the reference ITSELF (the one part of the hot path whose data can be
executed here: the tableau properties of `src/timesteppers/hdg_imex.py:668-1038` are plain numpy).

    python tests/golden/make_golden_tableaux.py [/root/reference]

The reference module cannot be imported (it imports firedrake), so the property bodies `nstages`, `_a_expl`,
`_a_impl`, `_b_expl`, `_b_impl`, `_c_expl` of every `IncompressibleEulerHDGIMEX*` class are cut out of the
source with `ast` and *executed* with numpy; the label is read from the `super().__init__(..., label=...)`
call.  Output: tests/golden/tableaux_v1.json.  `/root/reference` is only read here, never by the tests.
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROPS = ("nstages", "_a_expl", "_a_impl", "_b_expl", "_b_impl", "_c_expl")


def extract(path):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if not (isinstance(node, ast.ClassDef) and node.name.startswith("IncompressibleEulerHDGIMEX")):
            continue
        entry = {}
        for item in node.body:
            if isinstance(item, ast.FunctionDef) and item.name in PROPS:
                fn = ast.FunctionDef(name=item.name, args=item.args, body=item.body, decorator_list=[], returns=None,
                                     type_comment=None, type_params=[])
                mod = ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[]))
                ns = {"np": np}
                exec(compile(mod, path, "exec"), ns)
                val = ns[item.name](None)
                if val is None:  # the abstract base class
                    continue
                entry[item.name] = np.asarray(val, dtype=float).tolist() if item.name != "nstages" else int(val)
            if isinstance(item, ast.FunctionDef) and item.name == "__init__":
                for call in ast.walk(item):
                    if isinstance(call, ast.Call):
                        for kw in call.keywords:
                            if kw.arg == "label" and isinstance(kw.value, ast.Constant):
                                entry["label"] = kw.value.value
        if all(p in entry for p in PROPS):
            entry["lineno"] = node.lineno
            out[node.name] = entry
    return out


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    path = os.path.join(root, "src", "timesteppers", "hdg_imex.py")
    data = extract(path)
    with open(os.path.join(HERE, "tableaux_v1.json"), "w") as fh:
        json.dump({"source": "src/timesteppers/hdg_imex.py", "classes": data}, fh, indent=1)
    print("wrote tableaux_v1.json:", {k: (v["label"], v["nstages"]) for k, v in data.items()})
