#!/bin/bash
# GPU call 15 of round 2 (1 GPU): tabulated-Q* operator A/B, parity of the touched paths, smoke, bench
mkdir -p gpurun_out
T=r2o
timeout 500 python -m pytest tests/test_timesteppers_gpu.py tests/test_zz_cfl_gpu.py tests/test_golden.py tests/test_engine_flow_gpu.py tests/test_zz_more_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_part_${T}.log 2>&1; tail -3 gpurun_out/pytest_part_${T}.log
B="bench.py --steps 6 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
for v in "pre1:fimpl_pre=1" "pre0:fimpl_pre=0"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 200 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2o_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2o_")[1][:-5], round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), "| check", d["check"]["after_timed_region"]["l2_norm_velocity"], d["check"]["after_timed_region"]["l2_error_velocity"])
        print("  ", d["gpu_launches_per_step_by_kernel"])
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
echo done
