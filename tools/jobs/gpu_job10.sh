#!/bin/bash
# 2-GPU job: full GPU suite (includes the 1-vs-2 GPU parity via torchrun), 1-GPU bench, 2-GPU bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
timeout 900 python bench.py > gpurun_out/bench_r1l_1gpu.json 2> gpurun_out/bench_r1l_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_r1l_2gpu.json 2> gpurun_out/bench_r1l_2gpu.err
echo done
