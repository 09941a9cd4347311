"""Fully implicit (unsplit) stage solve, the replacement for the reference's monolithic
GMRES+MUMPS solve (`hdg_imex.py:600-620`; `hdg_implicit.py:153-186`).

The fully implicit operator is not cell-locally condensable: the momentum block
couples neighbouring cells directly.  The engine therefore runs a flexible outer Krylov iteration
on the monolithic (u, phi, lambda) operator, preconditioned by exactly the projection pair the
reference itself defines (tentative-velocity solve + statically condensed mixed Poisson,
`hdg_imex.py:572-599`), so that condensation and the trace solve appear inside the preconditioner.
"""

from __future__ import annotations

__all__ = ["MonolithicStage"]


class MonolithicStage:
    def __init__(self, timestepper):
        self.ts = timestepper

    def solve(self, Q_star, adt, rho, Q, p, lmbda):
        raise NotImplementedError(
            "the fully implicit (unsplit) stage solve is not implemented yet; use use_projection_method=True")
