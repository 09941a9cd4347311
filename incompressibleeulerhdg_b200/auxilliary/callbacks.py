"""Callbacks invoked at the end of every timestep, with the reference's interface
(`src/auxilliary/callbacks.py:11-85`): ``callback(Q, p, t, q_tracer=None)`` and ``callback.reset()``.

``AnimationCallback`` saves velocity, pressure, vorticity (and the tracer) to a ParaView collection.
The reference projects the weak vorticity  -eps : (grad tau (x) Q) dx + tau eps : (n (x) Q) ds
onto CG_{k+1} (`callbacks.py:44-69`); that is an L2 projection of curl Q in the weak sense.  Output is
off the hot path, so the vorticity here is the cell-wise curl of the DG velocity sampled at the cell
vertices -- the strong form of the same quantity, without the global CG mass solve.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .. import refelem as R
from .vtk import VTKFile

__all__ = ["Callback", "AnimationCallback"]


class Callback(ABC):
    """Abstract base class"""

    @abstractmethod
    def __call__(self, Q, p, t, q_tracer=None):
        """invoke the callback for velocity/pressure (and tracer) fields at time t"""

    @abstractmethod
    def reset(self):
        """reset the callback"""


def cell_vorticity_at_vertices(Q):
    """curl Q = d_x Q_y - d_y Q_x of a cell-wise velocity at the three vertices of every cell [nc, 3]"""
    space = Q.function_space()
    mesh = space.mesh()
    coef = Q.to_host()  # [nc, 2, nQ1]
    dtab = R.dubiner_grad(space.degree, R.REF_VERTS)  # [nQ1, 3, 2] reference gradients
    x = mesh.cell_xy
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)  # J[n, c, d] = d x_c / d xi_d
    Jinv = np.linalg.inv(J)  # Jinv[n, d, c]
    gref = np.einsum("nci,ivd->ncvd", coef, dtab)  # d Q_c / d xi_d at the vertices
    grad = np.einsum("ncvd,nde->ncve", gref, Jinv)  # d Q_c / d x_e
    return grad[:, 1, :, 0] - grad[:, 0, :, 1]


class AnimationCallback(Callback):
    """Save fields to disk (`callbacks.py:28-85`)"""

    def __init__(self, filename):
        self.filename = filename
        self.reset()

    def reset(self):
        """re-open the file"""
        self.outfile = VTKFile(self.filename, mode="w")

    def __call__(self, Q, p, t, q_tracer=None):
        fields = [Q, p, ("vorticity", cell_vorticity_at_vertices(Q))]
        if q_tracer is not None:
            fields.append(q_tracer)
        self.outfile.write(*fields, time=t)
