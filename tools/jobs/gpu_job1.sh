#!/bin/bash
# round-1 job (b): tests, solver-parameter probe, ncu --set full of the hottest kernels
# (reports stay on the box; only the CSV pages come back: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
python tools/probe_params.py --nx 512 > gpurun_out/probe_params_512.log 2>&1
BENCH="python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
$BENCH > gpurun_out/plain_b.log 2>&1 || exit 1
prof() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/prof_$1_source.csv
  ls -la /tmp/prof_$1.ncu-rep >> gpurun_out/ncu_$1.log
}
prof tent_sweep k_tent_sweep 400
prof fimpl k_fimpl 20
prof cg_spmv k_cg_spmv 40
prof ell_cheb k_ell_cheb 40
du -sh gpurun_out
echo done
