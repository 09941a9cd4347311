"""Common base of the HDG timesteppers, mirroring `src/timesteppers/common.py:15-144` of the
reference with the Firedrake/Slate/PETSc machinery replaced by the B200 engine.

The public interface is the reference's: ``IncompressibleEuler(mesh, degree, dt, label)``,
``get_timesteps``, ``label``, ``project_bdm`` and the abstract
``solve(Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False) -> (Q, p)``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from ..engine import HDGEngine
from ..functions import Function, FunctionSpace

__all__ = ["IncompressibleEuler", "AUTO_PARTITION"]

#: partition the mesh over the ranks of an initialised torch.distributed group (tests switch this off
#: to build a single-GPU comparison engine inside a multi-rank job)
AUTO_PARTITION = True


class IncompressibleEuler(ABC):
    """Abstract base class; owns the engine and the three function spaces
    [DG_{k+1}]^2, DG_k, DGT_k (`hdg_imex.py:65-69`)."""

    def __init__(self, mesh, degree, dt, label=None, device=0, tau=1.0, preconditioner="gtmg", cell_rank=None):
        """`mesh` is the complete mesh, exactly as in the reference.  When the process is one rank of
        an initialised ``torch.distributed`` group (one process per GPU, `torchrun`), the mesh is
        partitioned here -- the analogue of Firedrake distributing a mesh over COMM_WORLD -- with
        `cell_rank` (default: contiguous strips, :func:`partition.strip_partition`) and this rank
        keeps its local part only."""
        self._mesh = mesh
        self.degree = degree
        self._dt = dt
        self._label = label
        self.tau = tau
        self.local_mesh = None
        rank, world = 0, 1
        try:
            import torch.distributed as dist

            if AUTO_PARTITION and dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        if world > 1:
            from .. import partition

            if cell_rank is None:
                cell_rank = partition.strip_partition(mesh, world)
            self.local_mesh = partition.partition_mesh(mesh, cell_rank, rank, world)
            self.engine = HDGEngine(self.local_mesh, degree, tau=tau, device=device)
        else:
            self.engine = HDGEngine(mesh, degree, tau=tau, device=device)
        self._V_Q = FunctionSpace(self.engine, "Q")
        self._V_p = FunctionSpace(self.engine, "p")
        self._V_q = FunctionSpace(self.engine, "p")
        self._V_trace = FunctionSpace(self.engine, "trace")
        self._V = (self._V_Q, self._V_p, self._V_trace)
        # domain volume (`common.py:72-73`)
        self.domain_volume = mesh.volume
        # K1-K3 once: the mixed-Poisson operator does not depend on dt, Q* or t
        self.engine.setup_poisson()
        # trace preconditioner: "gtmg" = P1 coarse space + Chebyshev/facet-block-Jacobi smoothing
        # (the reference's firedrake.GTMGPC, hdg_imex.py:138-169); "jacobi" = facet-block-Jacobi only
        assert preconditioner in ("gtmg", "jacobi")
        self.preconditioner = preconditioner
        if preconditioner == "gtmg":
            try:
                self.engine.mg_setup(global_mesh=mesh if self.local_mesh is not None else None)
            except ValueError as exc:  # no nested P1 hierarchy for this mesh
                import warnings

                warnings.warn(f"falling back to facet-block-Jacobi CG: {exc}")
                self.preconditioner = "jacobi"

    def get_timesteps(self, t_final, warmup):
        """number of timesteps (`common.py:75-84`)"""
        nt = 1 if warmup else int(np.round(t_final / self._dt))
        assert warmup or (abs(nt * self._dt - t_final) < 1.0e-12), "dt must divide the final time"
        return nt

    @property
    def label(self):
        return self._label

    def project_bdm(self, Q: Function, out: Function | None = None) -> Function:
        """H(div)-conforming projection (`common.py:91-108`), kept in the cell-wise DG representation"""
        out = Function(self._V_Q) if out is None else out
        self.engine.project_bdm_dev(Q.data, out.data)
        return out

    @abstractmethod
    def solve(self, Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False):
        """propagate to T_final; returns the final velocity and pressure"""
