#!/bin/bash
# BASELINE.json configs through the driver mirror, one JSON-ish log per config (first GPU call of a round):
#   configs[0]  fully implicit k=1, 16x16, 10 steps                (also a golden vector, tests/test_golden.py)
#   configs[1]  fully implicit k=2, 256^2, time-dependent Psi(t)   (FGMRES on the monolithic system; never run
#               at this size in round 1 -- watch the outer iteration counts)
#   configs[3]  IMEX SSP2(3,3,2) k=3 with 2 Richardson iterations, reduced to 512^2 on one GPU
# usage: gpurun --timeout 900 -- bash tools/run_configs.sh [steps]
steps=${1:-2}
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  ( time timeout 400 python -m incompressibleeulerhdg_b200.driver "$@" --output none ) > "gpurun_out/config_${name}.log" 2>&1
  echo "rc=$?" >> "gpurun_out/config_${name}.log"
}
run 0_implicit_k1_nx16 --nx 16 --degree 1 --timestepper imex_implicit --dt 0.1 --tfinal 1.0 --kappa 0.5
# configs[1]: the driver's default dt = 0.04 is CFL 10 at nx = 256, where the FGMRES of the fully implicit stage stalls
# (profiles/r2/config_1_implicit_k2_nx256_dt0.04.log: 300 iterations per step, relative residual ~1e-1; the stepper
# now raises); the same config at CFL 0.32 and 1.3
run 1_implicit_k2_nx256_cfl032 --nx 256 --degree 2 --timestepper implicit --dt 0.00125 --tfinal $(python -c "print(0.00125*${steps})")
run 1_implicit_k2_nx256_cfl13 --nx 256 --degree 2 --timestepper implicit --dt 0.005 --tfinal $(python -c "print(0.005*${steps})")
run 3_imex_ssp2_k3_nx512 --nx 512 --degree 3 --timestepper imex_ssp2_332 --use_projection_method --richardson 2 \
    --dt 0.0005 --tfinal $(python -c "print(0.0005*${steps})")
echo done
