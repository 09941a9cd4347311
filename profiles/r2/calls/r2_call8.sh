#!/bin/bash
# GPU call 8 of round 2 (2 GPUs): where does the 1-vs-2 GPU check of the IMEX case hang?  (faulthandler stacks after
# 50 s, 15 s transport time-out), fused vs two-kernel exchange; then the tracer cases
mkdir -p gpurun_out
T=r2h
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542"
export HDG_DIST_DUMP_S=50 HDG_P2P_TIMEOUT_S=15
HDG_DIST_ONLY=imex_ssp2_k1 timeout 100 $TR tests/dist/run_dist_check.py > gpurun_out/dist_imex_${T}.jsonl 2> gpurun_out/dist_imex_${T}.err
echo "imex rc=$?"; cut -c1-250 gpurun_out/dist_imex_${T}.jsonl; grep -v "site-packages" gpurun_out/dist_imex_${T}.err | grep -B2 -A14 "most recent call first\|HDGError\|Error" | head -80
HDG_TUNING=p2p_fused=0 HDG_DIST_ONLY=imex_ssp2_k1 timeout 100 $TR tests/dist/run_dist_check.py > gpurun_out/dist_imex_unfused_${T}.jsonl 2> gpurun_out/dist_imex_unfused_${T}.err
echo "imex unfused rc=$?"; cut -c1-250 gpurun_out/dist_imex_unfused_${T}.jsonl
HDG_DIST_TRACER=1 HDG_DIST_ONLY=tracer timeout 150 $TR tests/dist/run_dist_check.py > gpurun_out/dist_tracer_${T}.jsonl 2> gpurun_out/dist_tracer_${T}.err
echo "tracer rc=$?"; cut -c1-300 gpurun_out/dist_tracer_${T}.jsonl; grep -v "site-packages" gpurun_out/dist_tracer_${T}.err | grep -B2 -A14 "most recent call first\|HDGError" | head -60
echo done
