#!/bin/bash
mkdir -p gpurun_out
HDG_CG_TRACE=1 timeout 300 python tools/debug_cg.py 1024 5 1 > gpurun_out/debug_cg_trace3.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
timeout 900 python bench.py > gpurun_out/bench_r1o_1gpu.json 2> gpurun_out/bench_r1o_1gpu.err
echo done
