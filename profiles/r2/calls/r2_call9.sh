#!/bin/bash
# GPU call 9 of round 2 (1 GPU): full suite, BASELINE configs through the driver, condensation sweep (configs[4]),
# ncu --set full of the dominant kernel and of k_condense<2>, default bench line with the compiled CPU baseline
mkdir -p gpurun_out
T=r2i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
tail -n 4 gpurun_out/pytest_gpu_${T}.log
timeout 500 python bench.py --steps 10 --warmup 6 > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err; echo "bench rc=$?"
HDG_TUNING=tent_mixed=1 timeout 300 python bench.py --steps 3 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 1 > gpurun_out/bench_${T}_mixed_highcfl.json 2> gpurun_out/bench_${T}_mixed_highcfl.err; echo "bench mixed rc=$?"
timeout 300 python tools/condense_bench.py --nc 1000000 10000000 --k 1 2 3 4 > gpurun_out/condense_bench_${T}.jsonl 2> gpurun_out/condense_bench_${T}.err; echo "condense rc=$?"
bash tools/run_configs.sh 2 > /dev/null 2>&1
for f in gpurun_out/config_*.log; do echo "== $f"; grep -v "^\s*$" $f | tail -n 14; done
bash tools/gpu_profile.sh ${T}_condense2 "k_condense" 1 -- python tools/condense_bench.py --nc 3000000 --k 2 --reps 2 > /dev/null 2>&1
bash tools/gpu_profile.sh ${T}_tent_sweep32 "k_tent_sweep32" 300 -- python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 > /dev/null 2>&1
bash tools/gpu_profile.sh ${T}_back_update "k_back_update" 1 -- python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 > /dev/null 2>&1
cat gpurun_out/ncu_${T}_*_summary.txt | head -80
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2i_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2i_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3),
              "| cold", d["cold_start"] and round(d["cold_start"]["value"],3), "| high", d["high_cfl"] and (d["high_cfl"].get("converged"), d["high_cfl"].get("ms_per_step"), d["high_cfl"].get("tentative_solver")))
        print("   breakdown", d["breakdown_ms_per_step"])
        print("   kernels", {k:(round(v.get("frac",0),3), v.get("launch_ms")) for k,v in d["other_kernels"].items() if "frac" in v})
        print("   cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_seconds_per_step"]))
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1200:])
for l in open("gpurun_out/condense_bench_r2i.jsonl"):
    try:
        d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,dict)}, {k:round(v.get("hbm_frac",0),3) for k,v in d.items() if isinstance(v,dict)})
    except Exception: pass
PY
du -sh gpurun_out; echo done
