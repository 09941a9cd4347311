#!/bin/bash
# 8-GPU strong-scaling A/B of round 2 (after tools/r2_first_call.sh has shown that the cell blocks work on one GPU):
#   gpurun --gpus 8 --timeout 900 -- bash tools/r2_scaling_call.sh
# default | cell blocks + 4 sweeps | the same with local Schur sweeps (no facet exchange between the sweeps; with the
# cell blocks the iteration count depends much less on the accuracy of the Schur solve)
mkdir -p gpurun_out
B="bench.py --steps 5 --warmup 6 --e2e-steps 2 --no-cpu-baseline --gpus 8"
i=0
for tune in "" tent_cellblock=1,tent_sweeps=4 tent_cellblock=1,tent_sweeps=4,tent_local_sweeps=1; do
  name=$( [ -z "$tune" ] && echo default || echo $tune | tr ',=' '__' )
  HDG_TUNING=$tune timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port $((29540 + i)) $B > gpurun_out/scale_r2a_strong8_${name}.json 2> gpurun_out/scale_r2a_strong8_${name}.err
  i=$((i + 1))
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_r2a_strong8_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("strong8_")[1][:-5], d["value"], "steps/s", d["iterations"]["tentative_bicgstab_per_solve"], "its", d["gpu_launches"], "launches")
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo done
