"""Observed convergence rates of the engine against the exact Taylor-Green solution Psi(t) Q_s, through the
driver (north_star: "identical observed convergence rates against the exact solution").  The expected values
are what the CPU oracle produces for exactly these runs; `tests/test_oracle_poisson.py::
test_timestepper_convergence_rates` is the CPU half (nx = 4 -> 8).  (The file name sorts last on purpose: the
rates need four complete IMEX runs, the cheap parity tests run first.)"""
import io

import numpy as np
import pytest

from incompressibleeulerhdg_b200 import driver

# (k, expected velocity rate, expected pressure rate) from nx = 8 -> 16: the values the CPU oracle produces for
# exactly this run (IMEX SSP2(3,3,2), dt = 0.0125, T = 0.05, errors against the interpolated exact solution
# Psi(t) Q_s, Psi(t)^2 p_s as `driver.py:365-380` computes them); asymptotically k+2 and k+1
OBSERVED_RATES = [(1, 2.775, 1.951), (2, 3.838, 2.990)]


@pytest.mark.gpu
@pytest.mark.parametrize("k,rate_Q,rate_p", OBSERVED_RATES)
def test_driver_observed_convergence_rates(k, rate_Q, rate_p):
    """north_star: "identical observed convergence rates against the exact solution Psi(t) Q_s" """
    errs = []
    for nx in (8, 16):
        res = driver.main(["--nx", str(nx), "--degree", str(k), "--dt", "0.0125", "--tfinal", "0.05",
                           "--use_projection_method", "--output", "none"], file=io.StringIO())
        errs.append((res["velocity_error"], res["pressure_error"]))
    got_Q = np.log2(errs[0][0] / errs[1][0])
    got_p = np.log2(errs[0][1] / errs[1][1])
    print(f"k={k} errors {errs} rates velocity {got_Q:.3f} pressure {got_p:.3f}")
    assert abs(got_Q - rate_Q) < 0.05 and abs(got_p - rate_p) < 0.05
    assert got_Q > k + 1.5 and got_p > k + 0.8
