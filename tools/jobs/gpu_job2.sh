#!/bin/bash
# 2-GPU job: single-GPU regression tests, then the 1-vs-2 GPU parity script
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist/run_dist_check.py > gpurun_out/dist_check_2.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist_check_2.log
echo done
