#!/bin/bash
# GPU call 1 of round 2 (trimmed tools/r2_first_call.sh; the default suite is green in GPUTEST_r01 at this HEAD)
mkdir -p gpurun_out
T=r2a
HDG_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_cellblock_gpu.py -m gpu -q -s \
    > gpurun_out/pytest_cellblock_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cellblock_${T}.log
HDG_TUNING=tent_cellblock=1,tent_sweeps=4,tent_flex=1,tent_fp32=1 timeout 400 python -m pytest tests -m gpu -q \
    > gpurun_out/pytest_gpu_${T}_knobs.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}_knobs.log
B="bench.py --steps 5 --warmup 6 --e2e-steps 2 --no-cpu-baseline"
for tune in none tent_cellblock=1 tent_cellblock=1,tent_sweeps=6 tent_cellblock=1,tent_sweeps=4 tent_cellblock=1,tent_sweeps=3 \
            tent_cellblock=1,tent_sweeps=4,tent_flex=1 tent_cellblock=1,tent_sweeps=4,tent_flex=1,tent_fp32=1 \
            tent_cellblock=1,tent_sweeps=3,tent_flex=1,tent_fp32=1; do
  name=$(echo $tune | tr ',=' '__')
  if [ $tune = none ]; then tune=""; fi
  HDG_TUNING=$tune timeout 200 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
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
        print(f.split("bench_r2a_")[1][:-5], round(d["value"],3), "steps/s", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2), "ms tentative",
              d["iterations"]["tentative_bicgstab_per_solve"], "its", d["other_kernels"].get("k_tent_sweep",{}).get("launch_ms"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/pytest_cellblock_${T}.log gpurun_out/pytest_gpu_${T}_knobs.log
echo done
