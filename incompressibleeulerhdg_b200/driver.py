"""Command-line driver with the options of the reference's ``src/driver.py`` (:24-178), running the
HDG timesteppers on the B200 engine::

    python -m incompressibleeulerhdg_b200.driver --nx 64 --degree 2 --timestepper imex_ssp2_332 \
        --use_projection_method --dt 0.01 --tfinal 0.1
    torchrun --nproc-per-node 8 -m incompressibleeulerhdg_b200.driver --nx 1024 ...   # one rank per GPU

Every option of the reference is accepted with the same name, choices and default.  Differences, all
forced by scope:

* ``--discretisation conforming|dg`` raises: only the HDG path is built.
* ``--timestepper implicit`` works (the reference passes an unexpected ``n_richardson`` keyword to
  ``IncompressibleEulerHDGImplicit`` and crashes).
* ``--test_pressure_solver`` times one condensed mixed-Poisson solve for a random velocity right-hand
  side drawn with ``PCG64(seed=123456789)`` (`driver.py:308-325`); the reference's call there no
  longer matches its own ``pressure_solve(key)`` signature.
* extra options ``--device`` and ``--output`` (directory for ``solution.pvd`` / ``evolution.pvd``;
  ``--output none`` skips the files).
"""

from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

from .auxilliary.callbacks import AnimationCallback
from .auxilliary.logging import log_summary
from .auxilliary.vtk import VTKFile
from .functions import Expression, Function
from .mesh import PeriodicSquareMesh, UnitDiskMesh, UnitSquareMesh
from .model_problems import DoubleLayerShearFlow, KelvinHelmholtz, TaylorGreen

TIMESTEPPERS = ["implicit", "imex_implicit", "imex_ars2_232", "imex_ars3_443", "imex_ssp2_332", "imex_ssp3_433"]


#: the reference's options (`driver.py:26-176`): (name, type, default, choices, what it selects).  Names, types,
#: defaults and choices are the reference's; `tests/test_driver.py` pins them against a copy of that list.
_VALUE_OPTIONS = [
    ("problem", str, "taylorgreen", ["taylorgreen", "kelvinhelmholtz", "shear"], "which model problem"),
    ("nx", int, 8, None, "cells per direction of the square meshes"),
    ("refinement", int, 2, None, "refinement level of the disk mesh"),
    ("degree", int, 1, None, "pressure degree k (velocity k+1, trace k)"),
    ("tfinal", float, 1.0, None, "end of the simulated interval"),
    ("kappa", float, 0.5, None, "decay rate of the Taylor-Green amplitude Psi(t)"),
    ("dt", float, 0.04, None, "time step"),
    ("discretisation", str, "hdg", ["conforming", "dg", "hdg"], "spatial discretisation (only hdg is built)"),
    ("richardson", int, 2, None, "Richardson iterations per IMEX stage"),
    ("flux", str, "upwind", ["upwind", "centered"], "numerical flux of the advection term"),
    ("timestepper", str, "imex_ssp2_332", TIMESTEPPERS, "time integrator"),
    ("forcing", str, "exponential", ["exponential", "constant"], "time dependence of the Taylor-Green forcing"),
]
_FLAG_OPTIONS = [
    ("use_projection_method", "split every implicit stage into tentative velocity + pressure correction"),
    ("test_pressure_solver", "time one condensed mixed-Poisson solve and exit"),
    ("warmup", "run a single time step only"),
    ("animation", "write the fields after every step to evolution.pvd"),
    ("tracer_advection", "advect the passive tracer sin(2 pi x) sin(2 pi y)"),
]


def build_parser() -> argparse.ArgumentParser:
    """argument parser with the reference's option set plus the engine-side ``--device`` / ``--output``"""
    parser = argparse.ArgumentParser("Mesh specifications and polynomial degree")
    for name, typ, default, choices, text in _VALUE_OPTIONS:
        parser.add_argument(f"--{name}", type=typ, default=default, choices=choices, help=f"{text} [{default}]")
    for name, text in _FLAG_OPTIONS:
        parser.add_argument(f"--{name}", action="store_true", default=False, help=text)
    parser.add_argument("--device", type=int, default=None, help="CUDA device [LOCAL_RANK, else 0]")
    parser.add_argument("--output", type=str, default=".", help="directory of the .pvd/.vtu files, 'none' to skip [.]")
    return parser


def build_mesh(args):
    """`driver.py:180-185`"""
    if args.problem == "taylorgreen":
        return UnitSquareMesh(args.nx, args.nx, quadrilateral=False)
    if args.problem == "shear":
        return PeriodicSquareMesh(args.nx, args.nx, L=2 * np.pi, quadrilateral=False)
    return UnitDiskMesh(refinement_level=args.refinement)


def timestepper_class(name: str):
    from . import timesteppers as T

    return {
        "implicit": T.IncompressibleEulerHDGImplicit,
        "imex_implicit": T.IncompressibleEulerHDGIMEXImplicit,
        "imex_ars2_232": T.IncompressibleEulerHDGIMEXARS2_232,
        "imex_ars3_443": T.IncompressibleEulerHDGIMEXARS3_443,
        "imex_ssp2_332": T.IncompressibleEulerHDGIMEXSSP2_332,
        "imex_ssp3_433": T.IncompressibleEulerHDGIMEXSSP3_433,
    }[name]


def build_timestepper(args, mesh, callbacks, device):
    """`driver.py:189-283`"""
    if args.discretisation != "hdg":
        raise RuntimeError(f"discretisation '{args.discretisation}' is outside the scope of the B200 engine "
                           "(only the HDG hybridisation path is built, SURVEY.md §2.1)")
    cls = timestepper_class(args.timestepper)
    kw = dict(flux=args.flux, use_projection_method=args.use_projection_method, callbacks=callbacks, device=device)
    if args.timestepper != "implicit":
        kw["n_richardson"] = args.richardson
    return cls(mesh, args.degree, args.dt, **kw)


def print_header(args, timestepper, file=None):
    """the banner and run summary the reference prints (`driver.py:285-306`), same labels in the same order"""
    per_problem = {"taylorgreen": [("mesh size", f"{args.nx} x {args.nx}"), ("forcing", args.forcing), ("kappa", args.kappa)],
                   "shear": [("mesh size", f"{args.nx} x {args.nx}")],
                   "kelvinhelmholtz": [("mesh refinement", args.refinement)]}[args.problem]
    rows = [("model problem", args.problem), *per_problem, ("polynomial degree", args.degree),
            ("final time", args.tfinal), ("timestep size", args.dt), ("discretisation", args.discretisation),
            ("numerical flux", args.flux), ("number of Richardson iterations", args.richardson),
            ("use projection method", args.use_projection_method), ("advect tracer", args.tracer_advection),
            ("timestepping method", timestepper.label)]
    title = "! timesteppers for incompressible Euler equations !"
    rule = "+" + "-" * (len(title) - 2) + "+"
    print("\n".join([rule, title, rule, ""] + [f"{label} = {value}" for label, value in rows] + [""]), file=file)


def test_pressure_solver(timestepper, file=None):
    """`driver.py:308-325`: solve the mixed-Poisson problem for b(w) = (f_Q, w), f_Q ~ N(0,1) nodal
    values; returns (seconds, iterations) of the second (warm) solve"""
    eng = timestepper.engine
    V_Q = timestepper._V_Q
    rng = np.random.Generator(np.random.PCG64(seed=123456789))
    nodal = rng.standard_normal((eng.nc, V_Q.nodes.shape[0], 2))
    f_Q = Function(V_Q, data=eng.upload(0, np.einsum("iq,nqc->nci", V_Q.Vinv, nodal)))
    b = Function(V_Q)
    eng.mass_dev(0, f_Q.data, b.data)
    Q, p, l = V_Q.zeros(), timestepper._V_p.zeros(), timestepper._V_trace.zeros()
    print("=== Testing pressure solver", file=file)
    print(file=file)
    eng.poisson_apply_dev(b.data, None, None, Q.data, p.data, l.data, rtol=1e-12, maxit=100000)
    for f in (Q, p, l):
        f.data.zero_()
    eng.synchronize()
    t_start = time.perf_counter()
    its = eng.poisson_apply_dev(b.data, None, None, Q.data, p.data, l.data, rtol=1e-12, maxit=100000)
    eng.synchronize()
    t_finish = time.perf_counter()
    print(f"    solve time           = {t_finish-t_start:12.4f} s", file=file)
    print(f"    number of iterations = {its}", file=file)
    return t_finish - t_start, its


def divergence(timestepper, Q):
    """L2 projection of div Q onto the pressure space (`driver.py:353-362`)"""
    eng = timestepper.engine
    divQ = Function(timestepper._V_p, name="divergence")
    eng.weak_divergence_dev(Q.data, divQ.data, scale=1.0, mode=0)
    eng.mass_dev(1, divQ.data, divQ.data, inverse=True)
    return divQ


def l2_norm(engine, f):
    return float(np.sqrt(engine.l2_inner_dev(f.space.kind, f.data, f.data)))


def main(argv=None, file=None):
    """run the driver; returns a dict with the final fields and error norms (the reference prints
    them; returning them as well makes the driver testable)"""
    args = build_parser().parse_args(argv)
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    rank = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # torchrun: one process per GPU
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        rank = dist.get_rank()
    if rank != 0 and file is None:
        file = open(os.devnull, "w")
    outdir = None if args.output == "none" else args.output
    if outdir is not None:
        os.makedirs(outdir, exist_ok=True)
    t_start = time.perf_counter()
    mesh = build_mesh(args)
    callbacks = None
    if args.animation and outdir is not None:
        callbacks = [AnimationCallback(os.path.join(outdir, f"evolution{'' if rank == 0 else rank}.pvd"))]
    timestepper = build_timestepper(args, mesh, callbacks, device)
    print_header(args, timestepper, file=file)
    result = {"args": args, "timestepper": timestepper}

    if args.test_pressure_solver:
        result["pressure_solver"] = test_pressure_solver(timestepper, file=file)
        return result

    if args.warmup:
        print("WARNING: performing a single timestep only!", file=file)
        print(file=file)

    if args.problem == "taylorgreen":
        model_problem = TaylorGreen(timestepper._V_Q, timestepper._V_p, args.forcing, args.kappa)
    elif args.problem == "shear":
        model_problem = DoubleLayerShearFlow(timestepper._V_Q, timestepper._V_p)
    else:
        model_problem = KelvinHelmholtz(timestepper._V_Q, timestepper._V_p)

    Q_0, p_0 = model_problem.initial_condition()
    q_0 = None
    if args.tracer_advection:  # `driver.py:340-342`
        q_0 = Expression(lambda x, y: np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y), 0)
    setup_s = time.perf_counter() - t_start
    Q, p = timestepper.solve(Q_0, p_0, q_0, model_problem.f_rhs(), args.tfinal, warmup=args.warmup)
    result.update(Q=Q, p=p, q_tracer=getattr(timestepper, "q_tracer", None))

    log_summary(file=file)
    # engine resources (no counterpart in the reference's output): host set-up time and device memory of this rank
    try:
        import torch

        free_b, total_b = torch.cuda.mem_get_info(device)
        result.update(setup_seconds=setup_s, device_memory_gb=(total_b - free_b) / 2**30)
        print(f"engine: {int(os.environ.get('WORLD_SIZE', '1'))} rank(s), {timestepper.engine.nc} local cells on rank 0, "
              f"set-up {setup_s:.1f} s, device memory in use {(total_b - free_b) / 2**30:.2f} GB", file=file)
    except Exception:  # reporting only
        pass

    if not args.warmup:
        eng = timestepper.engine
        Q.rename("velocity")
        p.rename("pressure")
        divQ = divergence(timestepper, Q)
        result["divergence_norm"] = l2_norm(eng, divQ)
        output_fields = [Q, p, divQ]
        exact_solution = model_problem.solution(args.tfinal)
        if exact_solution is not None:  # `driver.py:365-382`
            Q_exact, p_exact = exact_solution
            Q_exact.rename("velocity_exact")
            p_exact.rename("pressure_exact")
            Q_error = Function(timestepper._V_Q, name="velocity_error")
            eng.lincomb_dev(Q_error.data, [(1.0, Q.data), (-1.0, Q_exact.data)])
            p_error = Function(timestepper._V_p, name="pressure_error")
            eng.lincomb_dev(p_error.data, [(1.0, p.data), (-1.0, p_exact.data)])
            Q_error_nrm = l2_norm(eng, Q_error)
            p_error_nrm = l2_norm(eng, p_error)
            print(file=file)
            print(f"velocity error = {Q_error_nrm}", file=file)
            print(f"pressure error = {p_error_nrm}", file=file)
            print(file=file)
            result.update(velocity_error=Q_error_nrm, pressure_error=p_error_nrm)
            output_fields += [Q_exact, Q_error, p_exact, p_error]
        if outdir is not None:
            outfile = VTKFile(os.path.join(outdir, f"solution{'' if rank == 0 else rank}.pvd"))
            outfile.write(*output_fields)
    return result


if __name__ == "__main__":
    main()
    sys.exit(0)
