#!/usr/bin/env python
"""Per-kernel timings of the condensed mixed-Poisson path on one GPU (CUDA events, L2-sized inputs).

usage: python tools/microbench.py [--nx 1024] [--k 2] [--peaks]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.engine import HDGEngine  # noqa: E402
from incompressibleeulerhdg_b200.mesh import UnitSquareMesh  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def peaks():
    out = {}
    n = 1 << 28
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = timeit(lambda: y.copy_(x))
    out["hbm_copy_GBs"] = 2 * n * 8 / ms / 1e6
    m = 8192
    A = torch.randn(m, m, dtype=torch.float64, device="cuda")
    B = torch.randn(m, m, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: torch.matmul(A, B), n=5, warm=2)
    out["dgemm_TFLOPs"] = 2 * m ** 3 / ms / 1e9
    # plain FMA throughput (vector FP64 pipe)
    z = torch.randn(1 << 24, dtype=torch.float64, device="cuda")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1024)
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--peaks", action="store_true")
    ap.add_argument("--cg", type=int, default=200)
    args = ap.parse_args()
    res = {"nx": args.nx, "k": args.k}
    if args.peaks:
        res["peaks"] = peaks()
    t0 = time.time()
    m = UnitSquareMesh(args.nx, perturb=0.1)
    res["mesh_s"] = time.time() - t0
    t0 = time.time()
    eng = HDGEngine(m, args.k)
    eng.use_torch_stream()
    res["create_s"] = time.time() - t0
    k = args.k
    nQ, np_, b = 2 * eng.nQ1, eng.np_, eng.nl1
    NL = 3 * b
    nA = nQ + np_
    res["setup_ms"] = timeit(lambda: eng.setup_poisson(keep_local=True), n=3, warm=1)
    tm = eng.timers()
    Ru = torch.randn(nQ * m.nc, dtype=torch.float64, device="cuda")
    Rp = torch.randn(np_ * m.nc, dtype=torch.float64, device="cuda")
    Rp[: m.nc] -= Rp[: m.nc].mean()
    lam = torch.randn(b * m.nf, dtype=torch.float64, device="cuda")
    out_l = torch.empty_like(lam)
    Q, p = torch.empty_like(Ru), torch.empty_like(Rp)
    ms = timeit(lambda: eng.trace_spmv_dev(lam, out_l))
    nblocks = 5 * m.nf
    spmv_bytes = nblocks * b * b * 8 + nblocks * 4 + 2 * b * m.nf * 8
    res["spmv_ms"] = ms
    res["spmv_GBs"] = spmv_bytes / ms / 1e6
    ms = timeit(lambda: eng.forward_eliminate_dev(None, Rp, None, out_l))
    res["forward_p_ms"] = ms
    res["forward_p_GBs"] = (6 + np_ + NL) * 8 * m.nc / ms / 1e6
    ms = timeit(lambda: eng.forward_eliminate_dev(Ru, Rp, None, out_l))
    res["forward_full_ms"] = ms
    ms = timeit(lambda: eng.back_substitute_dev(None, Rp, lam, Q, p))
    res["back_ms"] = ms
    res["back_GBs"] = (6 + np_ + NL + nA) * 8 * m.nc / ms / 1e6
    ms = timeit(lambda: eng.back_substitute_dev(Ru, Rp, lam, Q, p))
    res["back_full_ms"] = ms
    res["back_full_GBs"] = (6 + nA + NL + nA) * 8 * m.nc / ms / 1e6
    # CG iterations: fixed count (rtol 0 never converges)
    eng.reset_timers()
    torch.cuda.synchronize()
    t0 = time.time()
    its = eng.poisson_apply_dev(None, Rp, None, Q, p, out_l, rtol=0.0, maxit=args.cg, check=False)
    torch.cuda.synchronize()
    wall = time.time() - t0
    tm = eng.timers()
    res["cg_iters"] = its
    res["cg_ms_per_iter"] = tm["trace_solve"][0] / max(its, 1)
    res["apply_wall_s"] = wall
    res["cg_bytes_per_iter_GB"] = (spmv_bytes + 12 * b * m.nf * 8) / 1e9
    res["cg_GBs"] = (spmv_bytes + 12 * b * m.nf * 8) / res["cg_ms_per_iter"] / 1e6
    # full solve to 1e-12 to see the iteration count
    if args.nx <= 512:
        its = eng.poisson_apply_dev(None, Rp, None, Q, p, out_l, rtol=1e-12, maxit=100000, check=False)
        res["iters_to_1e-12"] = its
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
