#!/bin/bash
# 2-GPU job: regression tests (incl. 1-vs-2 GPU parity), 1-GPU bench, 2-GPU bench strong + weak
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1d_1gpu.json 2> gpurun_out/bench_r1d_1gpu.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r1d_2gpu_strong.json 2> gpurun_out/bench_r1d_2gpu_strong.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --scaling weak > gpurun_out/bench_r1d_2gpu_weak.json 2> gpurun_out/bench_r1d_2gpu_weak.err
echo done
