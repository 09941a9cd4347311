#!/bin/bash
# final scaling run on one 8-GPU box: strong 4 / 8 and weak 8 (1 and 2 GPUs come from the 2-GPU job)
mkdir -p gpurun_out
B="bench.py --steps 5 --warmup 6 --e2e-steps 2 --no-cpu-baseline"
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    $B --gpus $N > gpurun_out/scale_r1m_strong$N.json 2> gpurun_out/scale_r1m_strong$N.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 \
  $B --gpus 8 --scaling weak > gpurun_out/scale_r1m_weak8.json 2> gpurun_out/scale_r1m_weak8.err
echo done
