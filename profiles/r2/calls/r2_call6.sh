#!/bin/bash
# GPU call 6 of round 2: ncu launch list of one mixed-precision step (where does the time go?)
mkdir -p gpurun_out
T=r2f
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
    --log-file gpurun_out/launches_${T}.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --cold-steps 0 --high-cfl-steps 0 \
    > gpurun_out/launches_${T}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${T}.csv > gpurun_out/launches_${T}.md 2>/dev/null
gzip -9 gpurun_out/launches_${T}.csv
head -60 gpurun_out/launches_${T}.md
echo done
