#!/usr/bin/env python
"""Generate tests/golden/golden_tracer_v1.npz from the CPU oracle (oracle/tracer.py, oracle/timesteppers.py).

    python tests/golden/make_golden_tracer.py

PARITY UNPINNED (same caveat as make_golden.py): outputs of the numpy/scipy restatement, not of the
reference itself.  Cases (the passive-tracer path):

  project_k2     L2 projection of a seeded discontinuous velocity onto [CG_3]^2  (`common.py:119-122`)
  advect_k2      M^-1 _tracer_advection(chi, q, u) for seeded q and the projected velocity (`common.py:123-129`)
  chorin_k2      driver.py's tracer sin(2 pi x) sin(2 pi y) advected by 2 Chorin steps  (`hdg_implicit.py:93-96,192-193`)
  imex_ssp2_k1   the same tracer through 1 step of SSP2(3,3,2)  (`hdg_imex.py:415-448`)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from oracle.hdg_oracle import HDGOracle  # noqa: E402
from oracle.timesteppers import ChorinOracle, IMEXOracle, TaylorGreenOracle  # noqa: E402
from oracle.tracer import TracerOracle  # noqa: E402

SEED = 123456789


def tracer0(x, y):
    """`driver.py:340-342`"""
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)


def cases():
    out = {}
    m = UnitSquareMesh(4, perturb=0.15)
    o = HDGOracle(m, 2)
    t = TracerOracle(o)
    rng = np.random.default_rng(SEED)
    Q = rng.standard_normal((m.nc, 2, o.nQ1))
    q = rng.standard_normal((m.nc, o.np_))
    U = t.project_cg(Q)
    out["project_k2/Q"], out["project_k2/U"] = Q, U
    out["advect_k2/q"], out["advect_k2/adv"] = q, t.advection(q, U)
    m = UnitSquareMesh(4, perturb=0.1)
    orc = ChorinOracle(m, 2, 0.02)
    orc.solve(TaylorGreenOracle("exponential", 0.5), 0.04, q_initial=tracer0)
    out["chorin_k2/q"] = orc.q_tracer
    m = UnitSquareMesh(5, perturb=0.1)
    orc = IMEXOracle(m, 1, 0.02, tableau="imex_ssp2_332", n_richardson=2)
    orc.solve(TaylorGreenOracle("exponential", 0.5), 0.02, q_initial=tracer0)
    out["imex_ssp2_k1/q"] = orc.q_tracer
    return out


if __name__ == "__main__":
    data = cases()
    path = os.path.join(HERE, "golden_tracer_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})
