#!/bin/bash
# GPU call 12 of round 2 (1 GPU): configs[1] at CFL 0.32 / 1.3 through the driver, ncu --set full of a regular
# (non-zero iterate) k_tent_sweep32 launch, register-variant A/B of the sweep kernel, dist-free re-check of the suite parts touched
mkdir -p gpurun_out
T=r2l
run() {  # name, args...
  local name=$1; shift
  ( time timeout 300 python -m incompressibleeulerhdg_b200.driver "$@" --output none ) > "gpurun_out/config_${name}.log" 2>&1
  echo "rc=$?" >> "gpurun_out/config_${name}.log"
}
run 1_implicit_k2_nx256_cfl032 --nx 256 --degree 2 --timestepper implicit --dt 0.00125 --tfinal 0.0025
run 1_implicit_k2_nx256_cfl13 --nx 256 --degree 2 --timestepper implicit --dt 0.005 --tfinal 0.01
for f in gpurun_out/config_1_*.log; do echo "== $f"; grep -v "^\s*$" $f | tail -n 16; done
bash tools/gpu_profile.sh ${T}_tent_sweep32 "k_tent_sweep32" 302 -- python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 > /dev/null 2>&1
cat gpurun_out/ncu_${T}_tent_sweep32_summary.txt | head -20
B="bench.py --steps 5 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
for v in "mb6:sweep_minblocks=6" "mb8:sweep_minblocks=8"; do
  name=${v%%:*}; tune=${v#*:}
  HDG_TUNING=$tune timeout 200 python $B > gpurun_out/bench_${T}_${name}.json 2> gpurun_out/bench_${T}_${name}.err
done
timeout 300 python -m pytest tests/test_timesteppers_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_part_${T}.log 2>&1; tail -3 gpurun_out/pytest_part_${T}.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2l_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2l_")[1][:-5], round(d["value"],3), "steps/s | tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "| roofline", d["roofline"]["kernel"][:16], round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo done
