#!/bin/bash
# 1-GPU job: tests, condensation microbenchmark (configs[4]), solver-parameter probe, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
python tools/condense_bench.py --nc 1000000 3000000 > gpurun_out/condense_bench_r1e.jsonl 2> gpurun_out/condense_bench_r1e.err
python tools/probe_params.py --nx 512 --cold --sweeps 4 5 6 8 10 --mg 1,1 1,2 2,2 1,3 2,1 > gpurun_out/probe_params_512.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1e_1gpu.json 2> gpurun_out/bench_r1e_1gpu.err
echo done
