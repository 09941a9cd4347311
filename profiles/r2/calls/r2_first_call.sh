#!/bin/bash
# First GPU call of round 2: everything round 1 left unmeasured, in one box (about 12 minutes).
#   gpurun --timeout 1500 -- bash tools/r2_first_call.sh
# 1. the full GPU parity suite of the default build (includes the tests written after the last GPU call)
# 2. the experimental cell-block advection preconditioner and flexible update: their own
#    tests, then the whole suite with the knobs on
# 3. bench A/B: default | cell blocks | cell blocks with 6 / 4 / 3 Schur sweeps | + flexible update (one JSON line each)
# 4. ncu launch list of the best candidate (share of every kernel in a step)
# Afterwards (separate calls, one kernel each): tools/gpu_profile.sh for k_tent_sweep, k_advblock_apply, k_fimpl.
mkdir -p gpurun_out
T=r2a
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
HDG_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_cellblock_gpu.py -m gpu -q -s \
    > gpurun_out/pytest_cellblock_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cellblock_${T}.log
# the whole parity suite once more with the experimental knobs on (every timestepper test then runs through them)
HDG_TUNING=tent_cellblock=1,tent_sweeps=4,tent_flex=1,tent_fp32=1 timeout 400 python -m pytest tests -m gpu -q \
    > gpurun_out/pytest_gpu_${T}_knobs.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}_knobs.log
B="bench.py --steps 5 --warmup 6 --e2e-steps 2"
timeout 240 python $B > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
for tune in tent_cellblock=1 tent_cellblock=1,tent_sweeps=6 tent_cellblock=1,tent_sweeps=4 tent_cellblock=1,tent_sweeps=3 \
            tent_cellblock=1,tent_sweeps=4,tent_flex=1 tent_cellblock=1,tent_sweeps=4,tent_flex=1,tent_fp32=1; do
  name=$(echo $tune | tr ',=' '__')
  HDG_TUNING=$tune timeout 200 python $B --no-cpu-baseline > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
HDG_TUNING=tent_cellblock=1,tent_sweeps=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/launches_${T}.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline \
    > gpurun_out/launches_${T}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${T}.csv > gpurun_out/launches_${T}.md 2>/dev/null
gzip -9 gpurun_out/launches_${T}.csv
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2a_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2a_")[1][:-5], d["value"], "steps/s", d["breakdown_ms_per_step"]["tentative_velocity_solve"], "ms tentative",
              d["iterations"]["tentative_bicgstab_per_solve"], "its", d["other_kernels"].get("k_tent_sweep"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo done
