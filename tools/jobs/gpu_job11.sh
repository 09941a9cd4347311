#!/bin/bash
# 1-GPU job: bench with the scaling-run flags, k_tent_sweep register-variant probe
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 6 --e2e-steps 2 --no-cpu-baseline > gpurun_out/scale_r1m_strong1.json 2> gpurun_out/scale_r1m_strong1.err
timeout 900 python tools/probe_params.py --nx 1024 --steps 3 --sweeps --mg --minblocks 5 6 8 5 6 8 > gpurun_out/probe_minblocks_r1m.log 2>&1
echo done
