"""Small helpers with the reference's interface (`auxilliary/utils.py:11-79`)."""

from __future__ import annotations

import numpy as np

__all__ = ["Averager", "gridspacing"]


class Averager:
    """running mean of a stream of numbers (used for Krylov iteration counts, hdg_imex.py:90-93)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self._n = 0
        self._mean = 0

    def update(self, x):
        self._n += 1
        self._mean += (x - self._mean) / self._n

    def __repr__(self):
        return f"{self.value} (averaged over {self.n_samples} samples)"

    @property
    def value(self):
        return self._mean

    @property
    def n_samples(self):
        return self._n


def gridspacing(mesh):
    """(h_min, h_max): shortest and longest edge of the mesh (utils.py:49-79)"""
    h = mesh.facet_length()
    return float(np.min(h)), float(np.max(h))
