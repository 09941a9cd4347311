#!/usr/bin/env python
"""Multi-GPU parity check, run as  torchrun --nproc-per-node N tests/dist/run_dist_check.py  (one
rank per GPU, NCCL).  Every rank also solves the complete problem on its own GPU (single-GPU
engine) and compares the owned part of the partitioned result with it: SURVEY.md §4 tier T4,
"1 vs N GPU results close to 1e-12 on the same mesh".  Prints one JSON line per case on rank 0 and
exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from incompressibleeulerhdg_b200 import multigrid, partition  # noqa: E402
from incompressibleeulerhdg_b200.engine import HDGEngine, broadcast_unique_id  # noqa: E402
from incompressibleeulerhdg_b200.mesh import UnitDiskMesh, UnitSquareMesh  # noqa: E402
from incompressibleeulerhdg_b200.model_problems import TaylorGreen  # noqa: E402
from incompressibleeulerhdg_b200.timesteppers import common as ts_common  # noqa: E402
from incompressibleeulerhdg_b200 import timesteppers as TS  # noqa: E402

TOL = 1e-10
failures = []


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def report(rank, name, errs, extra=None):
    worst = max(errs.values())
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(t.item() < TOL)
    if rank == 0:
        print(json.dumps({"case": name, "ok": ok, "max_rel_err_over_ranks": t.item(), "rank0": errs, **(extra or {})}),
              flush=True)
    if not ok:
        failures.append(name)


def poisson_case(rank, world, local, mesh, k, pc, thr, name, p2p=True):
    cr = partition.strip_partition(mesh, world)
    lm = partition.partition_mesh(mesh, cr, rank, world)
    ref = HDGEngine(mesh, k, device=local)
    ref.setup_poisson()
    eng = HDGEngine(lm, k, device=local)
    if not p2p:
        eng.p2p_enable(False)  # NCCL send/recv + all-reduce instead of the peer-memory transport
    eng.setup_poisson()
    if pc == "gtmg":
        H = multigrid.build_hierarchy(mesh, k)
        ref.mg_setup(hierarchy=H)
        eng.mg_setup(hierarchy=H, global_mesh=mesh, repl_threshold=thr)
    rng = np.random.default_rng(7)
    sQ, sp_, sl = ref.shapes()
    Ru, Rp, Rl = rng.standard_normal(sQ), rng.standard_normal(sp_), rng.standard_normal(sl)
    Q0, p0, l0, it0 = ref.poisson_apply_host(Ru, Rp, Rl, rtol=1e-13)
    cg, fg = lm.cells.local_gid, lm.facets.local_gid
    Q1, p1, l1, it1 = eng.poisson_apply_host(Ru[cg], Rp[cg], Rl[fg], rtol=1e-13)
    nco, nfo = lm.nc_owned, lm.nf_owned
    errs = {"Q": rel(Q1[:nco], Q0[cg[:nco]]), "p": rel(p1[:nco], p0[cg[:nco]]), "l": rel(l1[:nfo], l0[fg[:nfo]])}
    report(rank, name, errs, {"its_single": it0, "its_dist": it1, "repl": getattr(eng, "hierarchy", None) and
                              getattr(eng.hierarchy, "repl", None), "comm": eng.comm_stats(), "p2p": p2p,
                              "p2p_timeouts": eng.p2p_status()})


def timestepper_case(rank, world, local, mesh, k, cls, kwargs, dt, nt, name, local_sweeps=False):
    def run(auto):
        ts_common.AUTO_PARTITION = auto
        ts = getattr(TS, cls)(mesh, k, dt, device=local, krylov_rtol=1e-13, **kwargs)
        ts.engine.set_tentative_comm(local_sweeps)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q0, p0 = prob.initial_condition()
        Q, p = ts.solve(Q0, p0, None, prob.f_rhs(), nt * dt)
        return ts, Q.to_host(), p.to_host()

    _, Q0, p0 = run(False)
    ts, Q1, p1 = run(True)
    lm = ts.local_mesh
    cg, nco = lm.cells.local_gid, lm.nc_owned
    errs = {"Q": rel(Q1[:nco], Q0[cg[:nco]]), "p": rel(p1[:nco], p0[cg[:nco]])}
    report(rank, name, errs, {"comm": ts.engine.comm_stats(), "local_sweeps": local_sweeps,
                              "its_tentative": ts.niter_tentative.value, "p2p_timeouts": ts.engine.p2p_status()})


def tracer_case(rank, world, local, mesh, k, cls, kwargs, dt, nt, name):
    """passive tracer on the partitioned mesh vs the single-GPU run (CG-dof halo plan, owned-only dots).
    (HDG_DIST_TRACER=0 skips it)"""
    from incompressibleeulerhdg_b200.functions import Expression

    q0 = Expression(lambda x, y: np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y), 0)

    def run(auto):
        ts_common.AUTO_PARTITION = auto
        ts = getattr(TS, cls)(mesh, k, dt, device=local, krylov_rtol=1e-13, **kwargs)
        prob = TaylorGreen(ts._V_Q, ts._V_p, "exponential", 0.5)
        Q0, p0 = prob.initial_condition()
        Q, p = ts.solve(Q0, p0, q0, prob.f_rhs(), nt * dt)
        return ts, Q.to_host(), ts.q_tracer.to_host()

    _, Q0, q_single = run(False)
    ts, Q1, q_dist = run(True)
    lm = ts.local_mesh
    cg, nco = lm.cells.local_gid, lm.nc_owned
    errs = {"Q": rel(Q1[:nco], Q0[cg[:nco]]), "q": rel(q_dist[:nco], q_single[cg[:nco]])}
    report(rank, name, errs, {"comm": ts.engine.comm_stats(), "cg_projection_its": ts.niter_cg_projection.value,
                              "p2p_timeouts": ts.engine.p2p_status()})


def only(name):
    """HDG_DIST_ONLY=<substring>: run just the matching cases (debugging aid)"""
    pat = os.environ.get("HDG_DIST_ONLY")
    return pat is None or pat in name


def main():
    import faulthandler

    # a rank that is stuck prints its Python stack (HDG_DIST_DUMP_S seconds after start, default 400)
    faulthandler.dump_traceback_later(float(os.environ.get("HDG_DIST_DUMP_S", "400")), exit=False)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    m16 = UnitSquareMesh(16, perturb=0.1)
    if only("poisson_k2_jacobi"):
        poisson_case(rank, world, local, m16, 2, "jacobi", 0, "poisson_k2_jacobi")
    if only("poisson_k2_gtmg_replicated"):
        poisson_case(rank, world, local, m16, 2, "gtmg", 100000, "poisson_k2_gtmg_replicated")
    if only("poisson_k2_gtmg_distributed_levels"):
        poisson_case(rank, world, local, m16, 2, "gtmg", 100, "poisson_k2_gtmg_distributed_levels")
    if only("poisson_k2_gtmg_distributed_levels_nccl"):
        poisson_case(rank, world, local, m16, 2, "gtmg", 100, "poisson_k2_gtmg_distributed_levels_nccl", p2p=False)
    if only("poisson_k1_gtmg_distributed"):
        poisson_case(rank, world, local, UnitSquareMesh(12, perturb=0.1), 1, "gtmg", 60, "poisson_k1_gtmg_distributed")
    if only("poisson_k3_disk_jacobi"):
        poisson_case(rank, world, local, UnitDiskMesh(3), 3, "jacobi", 0, "poisson_k3_disk_jacobi")
    if only("chorin_k2"):
        timestepper_case(rank, world, local, m16, 2, "IncompressibleEulerHDGImplicit", {}, 0.01, 2, "chorin_k2")
    if only("chorin_k2_local_sweeps"):
        timestepper_case(rank, world, local, m16, 2, "IncompressibleEulerHDGImplicit", {}, 0.01, 2,
                         "chorin_k2_local_sweeps", local_sweeps=True)
    if only("fully_implicit_k1"):
        timestepper_case(rank, world, local, UnitSquareMesh(12, perturb=0.1), 1, "IncompressibleEulerHDGImplicit",
                         {"use_projection_method": False}, 0.02, 1, "fully_implicit_k1")
    if only("imex_ssp2_k1"):
        timestepper_case(rank, world, local, UnitSquareMesh(12, perturb=0.1), 1, "IncompressibleEulerHDGIMEXSSP2_332",
                         {"n_richardson": 2}, 0.01, 1, "imex_ssp2_k1")
    if os.environ.get("HDG_DIST_TRACER", "1") == "1":
        if only("chorin_k2_tracer"):
            tracer_case(rank, world, local, m16, 2, "IncompressibleEulerHDGImplicit", {}, 0.01, 2, "chorin_k2_tracer")
        if only("imex_ssp2_k1_tracer"):
            tracer_case(rank, world, local, UnitSquareMesh(12, perturb=0.1), 1, "IncompressibleEulerHDGIMEXSSP2_332",
                        {"n_richardson": 2}, 0.01, 1, "imex_ssp2_k1_tracer")
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        if rank == 0:
            print("FAILED:", failures, flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
