#!/bin/bash
# GPU call 24 of round 2 (1 GPU, the last 5 GPU-minutes of the round): parity of the shared-memory-factor Poisson kernels
# (csrc/hdg_poisson_s.cuh, knob poisson_lsmem) against the oracle and the register kernels, the configs[3] element test
# (IMEX SSP2 k = 3) with the knob on, and the A/B of the k = 3, 4 condensation / forward / back kernels at 10^6 cells
mkdir -p gpurun_out
T=r2x
timeout 110 python -m pytest tests/test_zz_lsmem_gpu.py -m gpu -q -x > gpurun_out/pytest_lsmem_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lsmem_${T}.log
tail -n 4 gpurun_out/pytest_lsmem_${T}.log
timeout 110 python tools/condense_bench.py --nc 1000000 --k 4 3 --lsmem --reps 3 > gpurun_out/condense_bench_${T}.jsonl 2> gpurun_out/condense_bench_${T}.err; echo "bench rc=$?"
cat gpurun_out/condense_bench_${T}.jsonl | cut -c1-900
HDG_TUNING=poisson_lsmem=7 timeout 70 python -m pytest tests/test_timesteppers_gpu.py -m gpu -q -x -k "config3" > gpurun_out/pytest_config3_lsmem_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_config3_lsmem_${T}.log
tail -n 4 gpurun_out/pytest_config3_lsmem_${T}.log
echo done
