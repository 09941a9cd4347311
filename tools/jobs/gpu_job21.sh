#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_engine_poisson_gpu.py tests/test_golden.py -m gpu -q --tb=short > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s.log
timeout 40 python tools/condense_bench.py --nc 1000000 --k 3 4 --reps 3 > gpurun_out/condense_bench_r1s.jsonl 2> gpurun_out/condense_bench_r1s.err
timeout 40 ncu --set full --clock-control none --import-source on -k regex:k_tracer_adv_t -c 1 -o gpurun_out/ncu_r1s_tracer_adv python tools/tracer_bench.py 256 2 > gpurun_out/ncu_r1s_tracer_adv.log 2>&1
timeout 40 ncu --set full --clock-control none --import-source on -k regex:k_cgp_cellop -s 3 -c 1 -o gpurun_out/ncu_r1s_cgp_cellop python tools/tracer_bench.py 256 2 > gpurun_out/ncu_r1s_cgp_cellop.log 2>&1
timeout 40 ncu --set full --clock-control none --import-source on -k regex:k_condense -s 1 -c 1 -o gpurun_out/ncu_r1s_condense2 python tools/condense_bench.py --nc 1000000 --k 2 --reps 1 > gpurun_out/ncu_r1s_condense2.log 2>&1
echo done
