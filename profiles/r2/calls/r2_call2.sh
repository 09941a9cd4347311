#!/bin/bash
# GPU call 2 of round 2: full GPU suite with the new defaults (cell blocks, scaled Schur complement, flexible update,
# FP32 sweeps, FGMRES fallback, CFL tests), then bench A/B of the initial-guess order / scaled Schur / sweeps
mkdir -p gpurun_out
T=r2b
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
tail -n 5 gpurun_out/pytest_gpu_${T}.log
grep -h "CFL\|fgmres" gpurun_out/pytest_gpu_${T}.log | head -20
B="bench.py --steps 8 --warmup 8 --no-cpu-baseline"
timeout 400 python $B > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
for v in "order1:--warm-order 1:" "order2:--warm-order 2:" "order4:--warm-order 4:" "noscaledx::tent_scaledx=0" "sweeps3::tent_sweeps=3" "sweeps3_order2:--warm-order 2:tent_sweeps=3"; do
  name=${v%%:*}; rest=${v#*:}; flags=${rest%%:*}; tune=${rest#*:}
  HDG_TUNING=$tune timeout 300 python $B $flags --high-cfl-steps 0 --cold-steps 0 > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2b_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3),
              "| cold", d["cold_start"] and round(d["cold_start"]["value"],3), "| high", d["high_cfl"] and (d["high_cfl"].get("converged"), d["high_cfl"].get("ms_per_step"), d["high_cfl"].get("tentative_solver")))
        print("   check", d["check"]["after_timed_region"])
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1500:])
PY
echo done
