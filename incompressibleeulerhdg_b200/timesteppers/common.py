"""Common base of the HDG timesteppers, mirroring `src/timesteppers/common.py:15-144` of the
reference with the Firedrake/Slate/PETSc machinery replaced by the B200 engine.

The public interface is the reference's: ``IncompressibleEuler(mesh, degree, dt, label)``,
``get_timesteps``, ``label``, ``project_bdm`` and the abstract
``solve(Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False) -> (Q, p)``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from ..auxilliary.logging import PerformanceLog
from ..auxilliary.utils import Averager
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
        self.q_tracer = None
        #: tolerance of the CG velocity projection (<r,z> relative; Firedrake's `project` default is a
        #: CG solve to rtol 1e-8 [FD-knowledge] -- tighter here so that results are reproducible to 1e-10)
        self.cg_projection_rtol = 1e-13
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

    # -- passive tracer (`common.py:110-129`) ----------------------------------------------------------
    def _tracer_initialise(self, q_initial):
        """`Function(self._V_q, name="tracer").interpolate(q_initial)` (`hdg_implicit.py:73-77`,
        `hdg_imex.py:523-527`) plus the engine-side set-up of the [CG_{k+1}]^2 velocity projection;
        returns None when no tracer is advected"""
        self.q_tracer = None
        if q_initial is None or q_initial is False:
            return None
        # partitioned meshes: CG-dof halo plan + owned-only reductions (1-vs-2 GPU parity 6e-16 on the tracer, Chorin and
        # IMEX: profiles/r2/dist_tracer_r2h_2gpu.jsonl)
        self.engine.tracer_setup(global_mesh=self._mesh if self.local_mesh is not None else None)
        self.q_tracer = self._V_q.interpolate(q_initial)
        self.q_tracer.rename("tracer")
        self.niter_cg_projection = Averager()
        return self.q_tracer

    def _project_onto_cg(self, Q: Function, out: Function) -> Function:
        """`Function(V_CG).project(u)` (`common.py:119-122`), kept in the cell-wise representation"""
        with PerformanceLog("cg_projection"):
            its = self.engine.project_cg_dev(Q.data, out.data, rtol=self.cg_projection_rtol)
        self.niter_cg_projection.update(its)
        return out

    def _tracer_advection(self, q: Function, u_cg: Function, out: Function, c0=0.0, acc=None, c1=1.0):
        """out = c0 acc + c1 M^-1 _tracer_advection(chi, q, u) for an already projected velocity"""
        self.engine.tracer_advection_dev(u_cg.data, q.data, out.data, c0=c0,
                                         acc=None if acc is None else acc.data, c1=c1)
        return out

    @abstractmethod
    def solve(self, Q_initial, p_initial, q_initial, f_rhs, T_final, warmup=False):
        """propagate to T_final; returns the final velocity and pressure"""
