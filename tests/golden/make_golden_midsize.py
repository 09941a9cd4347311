#!/usr/bin/env python
"""Mid-size golden vector: BASELINE.json configs[2] reduced to nx = 32 (2048 cells, 64 k velocity dofs), k = 2, two
Chorin steps at the bench's CFL (dt = 0.32 / nx) and two at CFL 3.2 -- computed ONCE with the CPU oracle (sparse-direct
solves, `oracle/timesteppers.py`) and committed, so that GPU parity is pinned above toy sizes without re-running the slow
oracle in the test suite.  The compiled CPU baseline (oracle/cpu_ref) is checked against the same vectors.

    python tests/golden/make_golden_midsize.py      # ~1 minute
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402
from oracle.timesteppers import ChorinOracle, TaylorGreenOracle  # noqa: E402

NX, K, NT = 32, 2, 2


def main():
    out = {}
    mesh = UnitSquareMesh(NX, perturb=0.1)
    for name, dt in (("cfl032", 0.32 / NX), ("cfl32", 3.2 / NX)):
        Q, p = ChorinOracle(mesh, K, dt).solve(TaylorGreenOracle("exponential", 0.5), NT * dt)
        out[f"chorin_k2_nx32_{name}/Q"] = Q
        out[f"chorin_k2_nx32_{name}/p"] = p
        out[f"chorin_k2_nx32_{name}/dt"] = np.array(dt)
        print(name, "dt", dt, "|Q|max", np.abs(Q).max(), "|p|max", np.abs(p).max())
    np.savez_compressed(os.path.join(HERE, "golden_midsize_v1.npz"), **out)


if __name__ == "__main__":
    main()
