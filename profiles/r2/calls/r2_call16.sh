#!/bin/bash
# GPU call 16 of round 2 (1 GPU): Schur sweeps without the local-facet switch (default), component-split operator
# k_fimpl_c (knob fimpl_split = 1 immediates / 2 constant-bank tables): parity of the touched paths, then A/B
mkdir -p gpurun_out
T=r2p
TESTS="tests/test_timesteppers_gpu.py tests/test_zz_cfl_gpu.py tests/test_golden.py tests/test_engine_flow_gpu.py tests/test_zz_more_parity_gpu.py"
timeout 400 python -m pytest $TESTS -m gpu -x -q > gpurun_out/pytest_part_${T}.log 2>&1; tail -3 gpurun_out/pytest_part_${T}.log
HDG_TUNING=fimpl_split=2 timeout 400 python -m pytest tests/test_timesteppers_gpu.py tests/test_zz_cfl_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_split_${T}.log 2>&1; tail -3 gpurun_out/pytest_split_${T}.log
B="bench.py --steps 6 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
for v in "base:" "mb6:sweep_minblocks=6" "mb8:sweep_minblocks=8" "split1:fimpl_split=1" "split2:fimpl_split=2" "split2mb6:fimpl_split=2,sweep_minblocks=6"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 200 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2p_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2p_")[1][:-5], round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"],
              "| fimpl ms", d["other_kernels"].get("k_fimpl",{}).get("launch_ms"), "| check", d["check"]["after_timed_region"]["l2_norm_velocity"], d["check"]["after_timed_region"]["l2_error_velocity"])
        print("  ", {k: round(v,1) for k, v in d["gpu_launches_per_step_by_kernel"].items()})
    except Exception as e:
        print(f, "unreadable:", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
echo done
