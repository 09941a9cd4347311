#!/bin/bash
# GPU call 29 of round 2 (1 GPU, the last seconds of the budget): configs[4] at 3 x 10^6 cells, k = 4 and 3, engine defaults
# beside the register / shared-factor A/B
mkdir -p gpurun_out
timeout 70 python tools/condense_bench.py --nc 3000000 --k 4 3 --lsmem --reps 3 > gpurun_out/condense_bench_r2B.jsonl 2> gpurun_out/condense_bench_r2B.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/condense_bench_r2B.jsonl
