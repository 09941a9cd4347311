#!/bin/bash
mkdir -p gpurun_out
HDG_CG_TRACE=1 timeout 200 python tools/debug_cg.py 1024 4 1 > gpurun_out/debug_cg_trace4.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1p_1gpu.json 2> gpurun_out/bench_r1p_1gpu.err
echo done
