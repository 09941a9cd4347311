#!/bin/bash
# GPU call 11 of round 2 (8 GPUs): strong-scaled bench at N = 8 and N = 4 (configs[2]), 1-vs-4 GPU parity of two cases
mkdir -p gpurun_out
T=r2k
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N"
  timeout 240 $TR bench.py --gpus $N --steps 10 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 > gpurun_out/bench_${T}_${N}gpu.json 2> gpurun_out/bench_${T}_${N}gpu.err; echo "bench$N rc=$?"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29559"
HDG_DIST_ONLY=chorin_k2 HDG_P2P_TIMEOUT_S=30 timeout 200 $TR tests/dist/run_dist_check.py > gpurun_out/dist_check_${T}_4gpu.jsonl 2> gpurun_out/dist_check_${T}_4gpu.err
echo "dist4 rc=$?"; cut -c1-200 gpurun_out/dist_check_${T}_4gpu.jsonl
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2k_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2k_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| trace ms", round(d["breakdown_ms_per_step"]["trace_solve"],2), "back", round(d["breakdown_ms_per_step"]["back_substitution"],3))
        print("   check", d["check"]["after_timed_region"])
        print("   comm", d["comm"])
        print("   kern", {k:(v.get("launch_ms")) for k,v in list(d["other_kernels"].items())+[("roof",d["roofline"])]}, d["setup"])
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1500:])
PY
echo done
