#!/bin/bash
# GPU call 5 of round 2: mixed-precision tentative solve (FP64 refinement around an FP32 BiCGStab) -- parity tests,
# traced bench at nx = 1024, inner-tolerance A/B, all-FP64 comparison
mkdir -p gpurun_out
T=r2e
timeout 400 python -m pytest tests/test_zz_mixed_gpu.py tests/test_zz_cfl_gpu.py -m gpu -x -q -s > gpurun_out/pytest_mixed_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mixed_${T}.log
tail -n 25 gpurun_out/pytest_mixed_${T}.log
B="bench.py --steps 4 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
HDG_TUNING=tent_trace=1 timeout 300 python $B > gpurun_out/bench_${T}_trace.json 2> gpurun_out/bench_${T}_trace.err
echo "trace rc=$?"; grep "hdg tent" gpurun_out/bench_${T}_trace.err | tail -45
for v in "tol40:tent_inner_tol=40" "tol60:tent_inner_tol=60" "fp64:tent_mixed=0" "sweeps3:tent_sweeps=3" "sweeps5:tent_sweeps=5"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 300 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
tail -n 4 gpurun_out/pytest_gpu_${T}.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2e_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2e_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3))
        print("   check", d["check"]["after_timed_region"], d["iterations"]["tentative_solver"], d["iterations"].get("tentative_mixed_precision"))
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1200:])
PY
echo done
