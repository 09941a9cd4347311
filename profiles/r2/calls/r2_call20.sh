#!/bin/bash
# GPU call 20 of round 2 (1 GPU): k_fimpl_t (rows of the operator staged by TMA bulk copies, knob fimpl_split=3):
# parity of the flow paths with the knob on, then A/B against the default with the in-situ kernel times
mkdir -p gpurun_out
T=r2t
HDG_TUNING=fimpl_split=3 timeout 300 python -m pytest tests/test_timesteppers_gpu.py tests/test_zz_cfl_gpu.py tests/test_golden.py tests/test_engine_flow_gpu.py -m gpu -x -q > gpurun_out/pytest_tma_${T}.log 2>&1; tail -3 gpurun_out/pytest_tma_${T}.log
B="bench.py --steps 6 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 --insitu-steps 1"
for v in "tma:fimpl_split=3" "base:"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 150 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err; echo "rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2t_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2t_")[1][:-5], round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"],
              "| check", d["check"]["after_timed_region"]["l2_norm_velocity"], d["check"]["after_timed_region"]["l2_error_velocity"], d["clocks"])
        i = d["insitu_kernel_times"]
        print("  ", {k: v for k, v in i.items() if k != "by_kernel"})
        for k, v in list(i["by_kernel"].items())[:12]:
            print(f"    {k:22s} {v['launches_per_step']:7.1f} x {v['us_per_launch']:8.1f} us = {v['ms_per_step']:7.2f} ms  {100*v['share']:5.1f} %")
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
echo done
