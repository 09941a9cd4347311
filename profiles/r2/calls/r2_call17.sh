#!/bin/bash
# GPU call 17 of round 2 (1 GPU): new defaults (fimpl_split=1, sweep_minblocks=6) with the in-situ kernel times
mkdir -p gpurun_out
T=r2q
timeout 300 python bench.py --steps 8 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 --insitu-steps 2 > gpurun_out/bench_${T}_insitu.json 2> gpurun_out/bench_${T}_insitu.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2q_insitu.json").read().strip().splitlines()[-1])
print(round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2), "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), d["clocks"])
i = d["insitu_kernel_times"]
print({k: v for k, v in i.items() if k != "by_kernel"})
for k, v in i["by_kernel"].items():
    print(f"  {k:22s} {v['launches_per_step']:7.1f} x {v['us_per_launch']:8.1f} us = {v['ms_per_step']:7.2f} ms  {100*v['share']:5.1f} %")
PY
tail -5 gpurun_out/bench_${T}_insitu.err
