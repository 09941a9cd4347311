#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_dist_gpu.py > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1f_1gpu.json 2> gpurun_out/bench_r1f_1gpu.err
echo done
