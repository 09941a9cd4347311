#!/bin/bash
# GPU call 3 of round 2: why does the nx = 1024 bench stop converging with the new defaults?  Traced tentative solves
# (HDG_TUNING=tent_trace=1) with the preconditioned-residual acceptance test, A/B without verification; CFL tests;
# the compiled CPU baseline on the box's host cores.
mkdir -p gpurun_out
T=r2c
nproc > gpurun_out/cpu_${T}.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/cpu_${T}.txt; free -g | head -2 >> gpurun_out/cpu_${T}.txt
timeout 300 python -m pytest tests/test_zz_cfl_gpu.py tests/test_timesteppers_gpu.py -m gpu -x -q -s > gpurun_out/pytest_cfl_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cfl_${T}.log
tail -n 3 gpurun_out/pytest_cfl_${T}.log
B="bench.py --steps 3 --warmup 5 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
HDG_TUNING=tent_trace=1 timeout 200 python $B > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
echo "default rc=$?"; grep "hdg tent" gpurun_out/bench_${T}_default.err | head -40
HDG_TUNING=tent_trace=1,tent_verify=0 timeout 200 python $B > gpurun_out/bench_${T}_noverify.json 2> gpurun_out/bench_${T}_noverify.err
echo "noverify rc=$?"; grep "hdg tent" gpurun_out/bench_${T}_noverify.err | head -40
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err
echo "reference rc=$?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            print("reference", d["value"], d["cpu_baseline"]["sample"], d["cpu_baseline"]["iterations_tentative_pressure"]); continue
        print(f.split("bench_r2c_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3))
        print("   check", d["check"]["after_timed_region"], d["iterations"]["tentative_solver"])
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1200:])
PY
echo done
