#!/bin/bash
# scaling run on one 8-GPU box: strong 1/2/4/8 (the named 1024^2 mesh) and weak 8
mkdir -p gpurun_out
B="bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline"
timeout 600 python $B > gpurun_out/scale_r1g_strong1.json 2> gpurun_out/scale_r1g_strong1.err
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    $B --gpus $N > gpurun_out/scale_r1g_strong$N.json 2> gpurun_out/scale_r1g_strong$N.err
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 \
  $B --gpus 8 --scaling weak > gpurun_out/scale_r1g_weak8.json 2> gpurun_out/scale_r1g_weak8.err
echo done
