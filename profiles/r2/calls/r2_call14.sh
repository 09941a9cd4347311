#!/bin/bash
# GPU call 14 of round 2 (8 GPUs): BASELINE configs[3] at its stated size -- IMEX SSP2(3,3,2), k = 3, 2048 x 2048
# squares (8 388 608 triangles), two projection-preconditioned Richardson iterations per stage, 8 GPUs, 2 steps at CFL 0.26
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562"
( time timeout 420 $TR -m incompressibleeulerhdg_b200.driver --nx 2048 --degree 3 --timestepper imex_ssp2_332 --use_projection_method \
    --richardson 2 --dt 0.000125 --tfinal 0.00025 --output none ) > gpurun_out/config_3_imex_ssp2_k3_nx2048_8gpu.log 2>&1
echo "rc=$?" >> gpurun_out/config_3_imex_ssp2_k3_nx2048_8gpu.log
grep -v "^\s*$" gpurun_out/config_3_imex_ssp2_k3_nx2048_8gpu.log | grep -v "its  \|average number\|------" | tail -n 22
free -g | head -2
echo done
