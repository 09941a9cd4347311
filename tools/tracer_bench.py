#!/usr/bin/env python
"""Time the passive-tracer kernels (CG velocity projection, advection) on one GPU; one JSON line.

  python tools/tracer_bench.py [nx] [k]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from incompressibleeulerhdg_b200.engine import HDGEngine  # noqa: E402
from incompressibleeulerhdg_b200.functions import Expression, FunctionSpace  # noqa: E402
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.perf_counter()
mesh = UnitSquareMesh(nx)
eng = HDGEngine(mesh, k)
t1 = time.perf_counter()
sp = eng.tracer_setup()
t2 = time.perf_counter()
V_Q, V_q = FunctionSpace(eng, "Q"), FunctionSpace(eng, "p")
S, Cc, pi = np.sin, np.cos, np.pi
Q = V_Q.interpolate(Expression(lambda x, y: (-Cc((x - .5) * pi) * S((y - .5) * pi), S((x - .5) * pi) * Cc((y - .5) * pi)), 1))
q = V_q.interpolate(Expression(lambda x, y: S(2 * pi * x) * S(2 * pi * y), 0))
U, out = eng.empty(0), eng.empty(1)
its = eng.project_cg_dev(Q.data, U)  # warm-up
eng.tracer_advection_dev(U, q.data, out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = 5
ev[0].record()
for _ in range(n):
    its = eng.project_cg_dev(Q.data, U)
ev[1].record()
for _ in range(n):
    eng.tracer_advection_dev(U, q.data, out)
ev[2].record()
torch.cuda.synchronize()
NP, NQ1 = (k + 1) * (k + 2) // 2, (k + 2) * (k + 3) // 2
adv_ms = ev[1].elapsed_time(ev[2]) / n
adv_bytes = mesh.nc * 8 * (6 + 2 * NQ1 + 4 * NP + NP) + mesh.nc * 4 * 6  # geometry, u, q + 3 neighbours, out, nbr maps
print(json.dumps({
    "nx": nx, "k": k, "cells": mesh.nc, "cg_dofs": sp.ndof, "host_setup_s": round(t2 - t1, 3),
    "mesh_engine_s": round(t1 - t0, 3), "project_cg_ms": ev[0].elapsed_time(ev[1]) / n, "pcg_iterations": its,
    "advection_ms": adv_ms, "advection_algorithmic_GBps": adv_bytes / adv_ms / 1e6,
}))
