#!/bin/bash
# GPU call 4 of round 2: acceptance of the tentative solve on the preconditioned residual with in-loop tightening
# (k_bi_resume); full GPU suite; traced bench at nx = 1024; default bench line
mkdir -p gpurun_out
T=r2d
B="bench.py --steps 4 --warmup 6 --no-cpu-baseline --cold-steps 2 --high-cfl-steps 0"
HDG_TUNING=tent_trace=1 timeout 300 python $B > gpurun_out/bench_${T}_trace.json 2> gpurun_out/bench_${T}_trace.err
echo "trace rc=$?"; grep "hdg tent" gpurun_out/bench_${T}_trace.err | head -60
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
tail -n 4 gpurun_out/pytest_gpu_${T}.log
timeout 600 python bench.py --steps 10 --warmup 6 > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
echo "default rc=$?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2d_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3),
              "| cold", d["cold_start"] and round(d["cold_start"]["value"],3), "| high", d["high_cfl"] and (d["high_cfl"].get("converged"), d["high_cfl"].get("ms_per_step"), d["high_cfl"].get("tentative_solver")))
        print("   check", d["check"]["after_timed_region"], d["iterations"]["tentative_solver"])
        print("   cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_seconds_per_step"]))
        print("   breakdown", d["breakdown_ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1200:])
PY
echo done
