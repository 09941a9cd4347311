#!/bin/bash
# GPU call 30 of round 2 (1 GPU, what is left of the budget): configs[4] at 10^7 cells, k = 4, engine defaults + A/B
mkdir -p gpurun_out
timeout 45 python tools/condense_bench.py --nc 10000000 --k 4 --lsmem --reps 2 > gpurun_out/condense_bench_r2C.jsonl 2> gpurun_out/condense_bench_r2C.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/condense_bench_r2C.jsonl
