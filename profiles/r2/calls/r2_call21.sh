#!/bin/bash
# GPU call 21 of round 2 (1 GPU): L1 prefetch of the neighbour coefficients (fimpl_split=4) and of the facet rows of the
# table as well (5) in k_fimpl_c, against the default (1), with the in-situ kernel times
mkdir -p gpurun_out
T=r2u
HDG_TUNING=fimpl_split=5 timeout 200 python -m pytest tests/test_zz_cfl_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_pf_${T}.log 2>&1; tail -2 gpurun_out/pytest_pf_${T}.log
B="bench.py --steps 6 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 --insitu-steps 1"
for v in "pf1:fimpl_split=4" "pf2:fimpl_split=5" "base:"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 150 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err; echo "rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2u_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2u_")[1][:-5], round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2),
              "| check", d["check"]["after_timed_region"]["l2_norm_velocity"], d["check"]["after_timed_region"]["l2_error_velocity"], d["clocks"])
        i = d["insitu_kernel_times"]
        print("  ", {k: v for k, v in i.items() if k != "by_kernel"})
        for k, v in list(i["by_kernel"].items())[:4]:
            print(f"    {k:22s} {v['launches_per_step']:7.1f} x {v['us_per_launch']:8.1f} us = {v['ms_per_step']:7.2f} ms  {100*v['share']:5.1f} %")
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
echo done
