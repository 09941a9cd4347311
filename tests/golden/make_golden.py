#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference cannot be imported here (Firedrake/PETSc absent), so
these vectors are outputs of `oracle/` (the numpy/scipy restatement of the reference forms), not of
the reference itself.  They pin the oracle against regressions and let the GPU parity tests compare
against committed numbers.  Every case is a reduced form of a BASELINE.json config:

  poisson_k{1,2}      one condensed mixed-Poisson solve (`hdg_imex.py:123-170`), seeded random residual
  chorin_k2           configs[2] reduced: Chorin projection, k=2, 2 steps  (`hdg_implicit.py:101-150`)
  implicit_k1_nx16    configs[0] literal: fully implicit, k=1, 16x16, stationary solution (kappa=0,
                      zero forcing), 10 steps dt=0.1 (`hdg_implicit.py:153-186`)
  imex_ssp2_k1        configs[3] reduced: SSP2(3,3,2) with projection-preconditioned Richardson, 1 step
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

SEED = 123456789  # the reference's own seed for its pressure-solver test (driver.py:309)


def poisson_case(k):
    m = UnitSquareMesh(4, perturb=0.15)
    o = HDGOracle(m, k)
    rng = np.random.default_rng(SEED + k)
    Ru = rng.standard_normal((m.nc, 2, o.nQ1))
    Rp = rng.standard_normal((m.nc, o.np_))
    Rl = rng.standard_normal((m.nf, k + 1))
    Q, p, l = o.solve_condensed(Ru, Rp, Rl)
    return dict(Ru=Ru, Rp=Rp, Rl=Rl, Q=Q, p=p, l=l)


class _Stationary(TaylorGreenOracle):
    """kappa = 0: Psi = 1, zero forcing (the reference's kappa==0 path crashes)"""

    def __init__(self):
        super().__init__("exponential", 0.0)


def cases():
    out = {}
    for k in (1, 2):
        for key, v in poisson_case(k).items():
            out[f"poisson_k{k}/{key}"] = v
    m = UnitSquareMesh(4, perturb=0.1)
    Q, p = ChorinOracle(m, 2, 0.02).solve(TaylorGreenOracle("exponential", 0.5), 0.04)
    out["chorin_k2/Q"], out["chorin_k2/p"] = Q, p
    m = UnitSquareMesh(16)
    orc = ChorinOracle(m, 1, 0.1, use_projection_method=False)
    Q, p = orc.solve(_Stationary(), 1.0)
    out["implicit_k1_nx16/Q"], out["implicit_k1_nx16/p"] = Q, p
    prob = _Stationary()
    out["implicit_k1_nx16/err_Q"] = np.array(orc.o.l2_error_Q(Q, prob.Q_stationary))
    m = UnitSquareMesh(5, perturb=0.1)
    Q, p = IMEXOracle(m, 1, 0.02, tableau="imex_ssp2_332", n_richardson=2).solve(
        TaylorGreenOracle("exponential", 0.5), 0.02)
    out["imex_ssp2_k1/Q"], out["imex_ssp2_k1/p"] = Q, p
    return out


if __name__ == "__main__":
    data = cases()
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})
