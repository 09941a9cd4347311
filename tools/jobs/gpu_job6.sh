#!/bin/bash
# 1-GPU job: tests, bench, ncu --set full of the two hottest tentative-solver kernels (CSV pages only)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1h_1gpu.json 2> gpurun_out/bench_r1h_1gpu.err
BENCH="python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
$BENCH > gpurun_out/plain_h.log 2>&1 || exit 1
prof() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_r1h_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv 2>/dev/null | head -c 2500000 > gpurun_out/prof_r1h_$1_source.csv
}
prof tent_sweep k_tent_sweep 403 3
prof fimpl k_fimpl 20 1
echo done
