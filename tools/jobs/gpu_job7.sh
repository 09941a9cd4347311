#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1i_1gpu.json 2> gpurun_out/bench_r1i_1gpu.err
echo done
