#!/bin/bash
# GPU call 13 of round 2 (8 GPUs): BASELINE configs[3] reduced to nx = 1024 -- IMEX SSP2(3,3,2), k = 3, two
# projection-preconditioned Richardson iterations per stage, through the driver on 8 GPUs (2 steps at CFL 0.26)
mkdir -p gpurun_out
T=r2m
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 420 $TR -m incompressibleeulerhdg_b200.driver --nx 1024 --degree 3 --timestepper imex_ssp2_332 --use_projection_method \
    --richardson 2 --dt 0.00025 --tfinal 0.0005 --output none ) > gpurun_out/config_3_imex_ssp2_k3_nx1024_8gpu.log 2>&1
echo "rc=$?" >> gpurun_out/config_3_imex_ssp2_k3_nx1024_8gpu.log
grep -v "^\s*$" gpurun_out/config_3_imex_ssp2_k3_nx1024_8gpu.log | tail -n 30
nvidia-smi --query-gpu=index,memory.used --format=csv | head -10
echo done
