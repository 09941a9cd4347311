#!/bin/bash
# GPU call 18 of round 2 (2 GPUs): new defaults (k_fimpl_c operator, sweeps without the local-facet switch, 80-register
# variant) on a partitioned mesh: 1-vs-2 GPU parity of the flow cases, strong-scaled bench
mkdir -p gpurun_out
T=r2r
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
for c in chorin imex_ssp2_k1; do
  HDG_DIST_ONLY=$c HDG_DIST_DUMP_S=80 HDG_P2P_TIMEOUT_S=30 timeout 150 $TR tests/dist/run_dist_check.py >> gpurun_out/dist_check_${T}_2gpu.jsonl 2>> gpurun_out/dist_check_${T}_2gpu.err
  echo "dist check $c rc=$?"
done
grep -c '"ok": true' gpurun_out/dist_check_${T}_2gpu.jsonl; grep '"ok"' gpurun_out/dist_check_${T}_2gpu.jsonl | cut -c1-160
grep -v "site-packages" gpurun_out/dist_check_${T}_2gpu.err | grep -A25 "most recent call first\|HDGError" | head -60
B="bench.py --gpus 2 --steps 8 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
timeout 200 $TR $B > gpurun_out/bench_${T}_2gpu.json 2> gpurun_out/bench_${T}_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
f = "gpurun_out/bench_r2r_2gpu.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
          "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| trace ms", round(d["breakdown_ms_per_step"]["trace_solve"],2))
    print("   check", d["check"]["after_timed_region"])
    print("   comm", d["comm"])
    print("   roof", d["roofline"]["kernel"][:20], d["roofline"]["launch_ms"], d["roofline"]["frac"])
except Exception as e:
    print(f, "unreadable:", e)
    print(open(f.replace(".json",".err")).read()[-1500:])
PY
echo done
