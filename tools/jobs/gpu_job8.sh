#!/bin/bash
# 2-GPU job: 1-vs-2 GPU parity with the peer-memory transport, then strong-scaling bench with P2P and with NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist/run_dist_check.py > gpurun_out/dist_check_j.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist_check_j.log
B="bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline"
timeout 600 $TR --master-port 29512 $B > gpurun_out/bench_r1j_2gpu_p2p.json 2> gpurun_out/bench_r1j_2gpu_p2p.err
HDG_P2P=0 timeout 600 $TR --master-port 29513 $B > gpurun_out/bench_r1j_2gpu_nccl.json 2> gpurun_out/bench_r1j_2gpu_nccl.err
echo done
