"""Model problems with the reference's interface (`src/model_problems.py:10-196`):
``TaylorGreen(V_Q, V_p, forcing, kappa)`` with ``initial_condition()``, ``f_rhs()`` and
``solution(t)``.  UFL expressions are replaced by :class:`Expression` callables of (x, y)."""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .functions import Expression, Function

__all__ = ["TaylorGreen", "KelvinHelmholtz", "DoubleLayerShearFlow"]


class ModelProblem(ABC):
    def __init__(self, V_Q, V_p):
        self.V_Q = V_Q
        self.V_p = V_p

    @abstractmethod
    def initial_condition(self):
        """(velocity expression, pressure expression)"""

    @abstractmethod
    def f_rhs(self):
        """callable t -> forcing expression"""

    def solution(self, t):
        return None


class TaylorGreen(ModelProblem):
    """Taylor-Green vortex with exact solution Psi(t) Q_s, Psi(t)^2 p_s (`model_problems.py:38-105`)"""

    def __init__(self, V_Q, V_p, forcing="exponential", kappa=0.5):
        super().__init__(V_Q, V_p)
        self.kappa = kappa
        assert forcing in ("exponential", "constant"), "Forcing must be 'constant' or 'exponential'"
        self.forcing = forcing
        S, Cc, pi = np.sin, np.cos, np.pi
        self.Q_stationary = Expression(
            lambda x, y: (-Cc((x - 0.5) * pi) * S((y - 0.5) * pi), S((x - 0.5) * pi) * Cc((y - 0.5) * pi)), 1)
        self.p_stationary = Expression(lambda x, y: (S((x - 0.5) * pi) ** 2 + S((y - 0.5) * pi) ** 2) / 2, 0)

    def initial_condition(self):
        return self.Q_stationary, self.p_stationary

    def f_rhs(self):
        """Psi'(t) Q_s.  (The reference returns the int 0 for kappa == 0, which its callers then
        try to call; a zero forcing expression is returned here instead.)"""
        if self.kappa == 0:
            return lambda t: 0.0 * self.Q_stationary
        if self.forcing == "exponential":
            return lambda t: (-self.kappa * np.exp(-self.kappa * t)) * self.Q_stationary
        return lambda t: (-self.kappa) * self.Q_stationary

    def solution(self, t):
        """interpolated exact solution at time t; the pressure is shifted by its integral as in the
        reference (`model_problems.py:104`: no division by the volume)"""
        psi = np.exp(-self.kappa * t) if self.forcing == "exponential" else 1.0 - self.kappa * t
        Q_exact = self.V_Q.interpolate(psi * self.Q_stationary)
        p_exact = self.V_p.interpolate(psi ** 2 * self.p_stationary)
        eng = self.V_p.engine
        one = self.V_p.interpolate(Expression(lambda x, y: 1.0 + 0 * x, 0))
        integral = eng.l2_inner_dev(1, p_exact.data, one.data)
        eng.lincomb_dev(p_exact.data, [(1.0, p_exact.data), (-integral, one.data)])
        return Q_exact, p_exact


class KelvinHelmholtz(ModelProblem):
    """rigid rotation inside r < 1/2, at rest outside, zero pressure, no forcing; meant for the
    unit disk mesh (`model_problems.py:108-131`)"""

    def __init__(self, V_Q, V_p):
        super().__init__(V_Q, V_p)
        r_max = 0.5

        def vel(x, y):
            inside = x * x + y * y < r_max ** 2
            return (np.where(inside, -y, 0.0), np.where(inside, x, 0.0))

        self.Q_stationary = Expression(vel, 1)
        self.p_stationary = Expression(lambda x, y: 0.0 * x, 0)
        self._zero = Expression(lambda x, y: (0.0 * x, 0.0 * x), 1)

    def initial_condition(self):
        return self.Q_stationary, self.p_stationary

    def f_rhs(self):
        return lambda t: self._zero


class DoubleLayerShearFlow(ModelProblem):
    """double shear layer of Guzman, Shu & Sequeira (2017) on the periodic square [0, 2 pi]^2
    (`model_problems.py:134-196`); the initial pressure is a 28-term sine series in y whose
    coefficients are oscillatory integrals of the shear profile"""

    def __init__(self, V_Q, V_p, rho=np.pi / 15, delta=0.05):
        super().__init__(V_Q, V_p)
        import scipy.integrate as integrate

        self.rho, self.delta = rho, delta

        def vel(x, y):
            u = np.where(y <= np.pi, np.tanh((y - np.pi / 2) / rho), np.tanh((1.5 * np.pi - y) / rho))
            return (u, delta * np.sin(x))

        def profile(z):
            lower = 1 - np.tanh((np.pi + 2 * z) / (4 * np.pi * rho)) ** 2
            upper = -1 + np.tanh((np.pi - 2 * z) / (4 * np.pi * rho)) ** 2
            return np.where(z <= 0.0, lower, upper) / (np.pi ** 2 * rho)

        kmax = 28
        modes = 2 * np.arange(kmax) + 1
        coef = np.array([
            integrate.quad(profile, -np.pi, np.pi, weight="sin", wvar=int(n), epsabs=1e-12, epsrel=1e-12)[0] / (1 + n ** 2)
            for n in modes])

        def pres(x, y):
            series = sum(c * np.sin(n * (y - np.pi)) for c, n in zip(coef, modes))
            return delta * np.cos(x) * series

        self.Q_initial = Expression(vel, 1)
        self.p_initial = Expression(pres, 0)
        self._zero = Expression(lambda x, y: (0.0 * x, 0.0 * x), 1)

    def initial_condition(self):
        return self.Q_initial, self.p_initial

    def f_rhs(self):
        return lambda t: self._zero
