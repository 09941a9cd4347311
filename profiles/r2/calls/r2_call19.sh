#!/bin/bash
# GPU call 19 of round 2 (1 GPU): ncu --set full of the two dominant kernels of the tentative-velocity iteration as shipped
# (k_fimpl_c, k_tent_sweep32 without the local-facet switch), one launch each inside a real step
mkdir -p gpurun_out
T=r2s
B="python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 --insitu-steps 0"
bash tools/gpu_profile.sh ${T}_fimpl_c "k_fimpl_c" 40 -- $B > /dev/null 2>&1
bash tools/gpu_profile.sh ${T}_tent_sweep32 "k_tent_sweep32" 300 -- $B > /dev/null 2>&1
cat gpurun_out/ncu_${T}_*_summary.txt | head -60
for k in fimpl_c tent_sweep32; do
  grep -E "Eligible Warps|Issued Warp|No Eligible|Executed Ipc|Issue Slots Busy|Theoretical Occupancy|Achieved Occupancy|Registers Per|L1/TEX Hit|L2 Hit|DRAM Throughput|Memory Throughput|Duration" gpurun_out/ncu_${T}_${k}_details.txt | head -24
  grep -E "stall|Stall" gpurun_out/ncu_${T}_${k}_details.txt | head -8
done
du -sh gpurun_out; echo done
