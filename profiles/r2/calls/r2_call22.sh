#!/bin/bash
# GPU call 22 of round 2 (1 GPU): final state -- full GPU suite, smoke, the driver's bench command, ncu launch list
mkdir -p gpurun_out
T=r2v
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}.log
tail -n 4 gpurun_out/pytest_gpu_${T}.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
f = "gpurun_out/bench_r2v_1gpu.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
          "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"],
          "| cold", d["cold_start"] and round(d["cold_start"]["value"],3), d["clocks"])
    print("   breakdown", d["breakdown_ms_per_step"])
    print("   kernels", {k:(round(v.get("frac",0),3), v.get("launch_ms"), v.get("insitu_launch_ms")) for k,v in d["other_kernels"].items() if "frac" in v})
    print("   high", [(h["dt"], h.get("converged"), h.get("ms_per_step"), h.get("tentative_solver")) for h in [d["high_cfl"]] + d["high_cfl_more"] if h])
    print("   cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_seconds_per_step"]))
    print("   check", d["check"])
    i = d["insitu_kernel_times"]
    print("  ", {k: v for k, v in i.items() if k != "by_kernel"})
    for k, v in list(i["by_kernel"].items())[:16]:
        print(f"    {k:22s} {v['launches_per_step']:7.1f} x {v['us_per_launch']:8.1f} us = {v['ms_per_step']:7.2f} ms  {100*v['share']:5.1f} %")
except Exception as e:
    print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${T}.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 --insitu-steps 0 > gpurun_out/launches_${T}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${T}.csv > gpurun_out/launches_${T}.md 2>/dev/null; head -16 gpurun_out/launches_${T}.md
gzip -9 gpurun_out/launches_${T}.csv
du -sh gpurun_out; echo done
