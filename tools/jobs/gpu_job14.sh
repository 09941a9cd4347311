#!/bin/bash
# 2-GPU job: full GPU suite (incl. 1-vs-2 GPU parity), default bench on 1 GPU, 2-GPU bench, sweep-variant probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n.log
timeout 900 python bench.py > gpurun_out/bench_r1n_1gpu.json 2> gpurun_out/bench_r1n_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_r1n_2gpu.json 2> gpurun_out/bench_r1n_2gpu.err
timeout 600 python tools/probe_params.py --nx 1024 --steps 3 --sweeps --mg --minblocks 5 6 8 5 6 8 > gpurun_out/probe_minblocks_r1n.log 2>&1
echo done
