#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_tracer_gpu.py tests/test_driver.py -m gpu -q -s --tb=short > gpurun_out/pytest_new_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_q.log
timeout 45 python tools/tracer_bench.py 512 2 > gpurun_out/tracer_bench_r1q.json 2> gpurun_out/tracer_bench_r1q.err
echo done
