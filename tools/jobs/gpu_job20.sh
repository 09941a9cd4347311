#!/bin/bash
mkdir -p gpurun_out
timeout 210 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r.log
timeout 45 python tools/condense_bench.py --nc 1000000 --k 1 2 3 4 --reps 3 > gpurun_out/condense_bench_r1r.jsonl 2> gpurun_out/condense_bench_r1r.err
timeout 30 python tools/tracer_bench.py 512 2 > gpurun_out/tracer_bench_r1r.json 2> gpurun_out/tracer_bench_r1r.err
echo done
