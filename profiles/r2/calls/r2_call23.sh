#!/bin/bash
# GPU call 23 of round 2 (1 GPU): the part of the GPU suite behind the (too strict) iteration-count assertion that
# stopped call 22, then the bench line with the dominant kernel selected by the in-situ shares
mkdir -p gpurun_out
T=r2w
timeout 400 python -m pytest tests/test_zz_cellblock_gpu.py tests/test_zz_cfl_gpu.py tests/test_zz_mixed_gpu.py tests/test_zz_more_parity_gpu.py tests/test_zz_rates_gpu.py -m gpu -q > gpurun_out/pytest_gpu_${T}_tail.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${T}_tail.log
tail -n 6 gpurun_out/pytest_gpu_${T}_tail.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
f = "gpurun_out/bench_r2w_1gpu.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
          "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"], d["roofline"].get("share_of_step"), d["roofline"].get("traffic"),
          "| cold", d["cold_start"] and round(d["cold_start"]["value"],3), d["clocks"])
    print("   kernels", {k:(round(v.get("frac",0),3), v.get("launch_ms"), v.get("share_of_step")) for k,v in d["other_kernels"].items() if "frac" in v})
    print("   cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_seconds_per_step"]))
except Exception as e:
    print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
echo done
