#!/bin/bash
# GPU call 7 of round 2 (2 GPUs): 1-vs-2 GPU parity incl. the tracer path, strong-scaled bench with the comm probe,
# fused vs two-kernel halo exchange
mkdir -p gpurun_out
T=r2g
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
HDG_DIST_TRACER=1 timeout 600 $TR tests/dist/run_dist_check.py > gpurun_out/dist_check_${T}_2gpu.jsonl 2> gpurun_out/dist_check_${T}_2gpu.err
echo "dist check rc=$?"; grep -c '"ok": true' gpurun_out/dist_check_${T}_2gpu.jsonl; grep '"ok": false' gpurun_out/dist_check_${T}_2gpu.jsonl | cut -c1-300; tail -5 gpurun_out/dist_check_${T}_2gpu.err
B="bench.py --gpus 2 --steps 5 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0"
timeout 400 $TR $B > gpurun_out/bench_${T}_2gpu.json 2> gpurun_out/bench_${T}_2gpu.err; echo "bench2 rc=$?"
HDG_TUNING=p2p_fused=0 timeout 400 $TR $B > gpurun_out/bench_${T}_2gpu_unfused.json 2> gpurun_out/bench_${T}_2gpu_unfused.err; echo "bench2 unfused rc=$?"
timeout 300 python bench.py --steps 5 --warmup 6 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_r2g_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_r2g_")[1][:-5], round(d["value"],3), "steps/s | e2e", round(d["e2e"]["value"],3), "| tent ms", round(d["breakdown_ms_per_step"]["tentative_velocity_solve"],2),
              "its", round(d["iterations"]["tentative_bicgstab_per_solve"],2), "cg", round(d["iterations"]["trace_cg_per_solve"],2), "| trace ms", round(d["breakdown_ms_per_step"]["trace_solve"],2), "back", round(d["breakdown_ms_per_step"]["back_substitution"],3))
        print("   check", d["check"]["after_timed_region"])
        print("   comm", d["comm"])
    except Exception as e:
        print(f, "unreadable:", e)
        print(open(f.replace(".json",".err")).read()[-1500:])
PY
echo done
