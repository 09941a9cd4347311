#!/usr/bin/env python
"""BASELINE.json configs[4]: batched condensation / back-substitution microbenchmark, k = 1..4,
10^6 .. 10^7 independent random-affine triangles (no two cells identical).

Per (k, nc) it reports CUDA-event times of k_condense (K1+K2: local operators + Schur complement),
k_assemble (K3), k_forward (a3) and k_back (a6), each with
  * achieved HBM GB/s on the algorithmic bytes of DESIGN.md section 4 and the fraction of the measured
    copy bandwidth (MEASURED_PEAKS.json), and
  * for the condensation, the "generic-route" FLOP rate (2/3 nA^3 + 2 nA^2 nl + 2 nl^2 nA per cell,
    SURVEY.md section 8 table: what Slate's dense LU executes) next to the measured FP64 FMA peak.  The
    engine's closed-form condensation executes far fewer flops than the generic route, so this
    equivalent rate may exceed the peak; the executed-instruction count comes from ncu.

usage: python tools/condense_bench.py [--nc 1000000 3000000] [--k 1 2 3 4]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.engine import HDGEngine  # noqa: E402
from incompressibleeulerhdg_b200.mesh import RandomAffineCells  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, nargs="+", default=[1_000_000, 3_000_000])
    ap.add_argument("--k", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lsmem", action="store_true",
                    help="k >= 3: also time the kernels with the Cholesky factor in shared memory "
                         "(csrc/hdg_poisson_s.cuh, hdg_set_tuning poisson_lsmem=7) and with it in registers "
                         "(poisson_lsmem=0) and report both beside the engine's defaults")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = float(peaks["hbm_gbs"])
    fp64 = None
    for nc in args.nc:
        mesh = RandomAffineCells(nc)
        for k in args.k:
            eng = HDGEngine(mesh, k)
            if fp64 is None:
                fp64 = eng.measure_fp64_peak()
                print(json.dumps({"fp64_fma_peak_tflops_measured": fp64, "hbm_gbs_measured": hbm}), flush=True)
            nQ, np_, b = 2 * eng.nQ1, eng.np_, eng.nl1
            nA, nl = nQ + np_, 3 * b
            # K >= 3 has two condensation kernels (hdg_set_tuning "condense_rows"): time the optional row-loop
            # kernel first, then the default unrolled thread-per-cell kernel whose numbers go into the line
            unrolled_ms = None
            if k >= 3:
                eng.set_tuning("condense_rows", 1)
                eng.setup_poisson(keep_local=True)
                eng.reset_timers()
                for _ in range(args.reps):
                    eng.setup_poisson(keep_local=True)
                eng.synchronize()
                tm0 = eng.timers()
                unrolled_ms = tm0["condense"][0] / tm0["condense"][1]
                eng.set_tuning("condense_rows", 0)
            eng.setup_poisson(keep_local=True)  # warm-up (allocations)
            eng.reset_timers()
            for _ in range(args.reps):
                eng.setup_poisson(keep_local=True)
            Ru = torch.randn(nQ * nc, dtype=torch.float64, device="cuda")
            Rp = torch.randn(np_ * nc, dtype=torch.float64, device="cuda")
            lam = torch.randn(b * mesh.nf, dtype=torch.float64, device="cuda")
            out_l = torch.empty_like(lam)
            Q, p = torch.empty_like(Ru), torch.empty_like(Rp)
            for _ in range(args.reps + 1):
                eng.forward_eliminate_dev(Ru, Rp, None, out_l)
                eng.back_substitute_dev(Ru, Rp, lam, Q, p)
            eng.synchronize()
            tm = eng.timers()
            t_cond = tm["condense"][0] / tm["condense"][1]
            t_asm = tm["assemble"][0] / tm["assemble"][1]
            t_fwd = tm["forward_elimination"][0] / tm["forward_elimination"][1]
            t_back = tm["back_substitution"][0] / tm["back_substitution"][1]
            lsmem = None
            if k >= 3 and args.lsmem:
                lsmem = {}
                for label, mask in (("register_kernels", 0), ("shared_factor_kernels", 7)):
                    eng.set_tuning("poisson_lsmem", mask)
                    eng.setup_poisson(keep_local=True)
                    eng.forward_eliminate_dev(Ru, Rp, None, out_l)
                    eng.back_substitute_dev(Ru, Rp, lam, Q, p)
                    eng.reset_timers()
                    for _ in range(args.reps):
                        eng.setup_poisson(keep_local=True)
                        eng.forward_eliminate_dev(Ru, Rp, None, out_l)
                        eng.back_substitute_dev(Ru, Rp, lam, Q, p)
                    eng.synchronize()
                    tl = eng.timers()
                    lsmem[label] = {name: tl[key][0] / tl[key][1] for name, key in
                                    (("condense_ms", "condense"), ("forward_ms", "forward_elimination"),
                                     ("back_ms", "back_substitution"))}
                eng.set_tuning("poisson_lsmem", -1)
            flops = (2.0 / 3.0) * nA ** 3 + 2.0 * nA * nA * nl + 2.0 * nl * nl * nA
            by_cond = (6 + 3 + nl * nl) * 8
            by_fwd = (6 + 3 + nA + nl) * 8 + 2 * 3 * b * 8  # + the facet gather/write of k_trace_rhs
            by_back = (6 + 3 + nA + nl + nA) * 8
            gbs = lambda by, ms: by * nc / ms / 1e6
            res = {
                "k": k, "nc": nc, "condense_ms": t_cond, "assemble_ms": t_asm, "forward_ms": t_fwd, "back_ms": t_back,
                "condense_GBs": gbs(by_cond, t_cond), "condense_hbm_frac": gbs(by_cond, t_cond) / hbm,
                "condense_generic_TFLOPs": flops * nc / t_cond / 1e9,
                "condense_generic_frac_of_fp64_peak": flops * nc / t_cond / 1e9 / fp64,
                "forward_GBs": gbs(by_fwd, t_fwd), "forward_hbm_frac": gbs(by_fwd, t_fwd) / hbm,
                "back_GBs": gbs(by_back, t_back), "back_hbm_frac": gbs(by_back, t_back) / hbm,
                "cells_per_s_condense": nc / t_cond * 1e3,
                "condense_rows_variant_ms": unrolled_ms,
            }
            if lsmem:  # A/B of hdg_set_tuning("poisson_lsmem", 0 / 7); the line above is the engine's default (k = 3: 5, k = 4: 7)
                for v in lsmem.values():
                    v.update(condense_hbm_frac=gbs(by_cond, v["condense_ms"]) / hbm,
                             forward_hbm_frac=gbs(by_fwd, v["forward_ms"]) / hbm,
                             back_hbm_frac=gbs(by_back, v["back_ms"]) / hbm)
                res["poisson_lsmem_ab"] = lsmem
            print(json.dumps(res), flush=True)
            del eng, Ru, Rp, lam, out_l, Q, p
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
