#!/bin/bash
# GPU call 25/26 of round 2 (1 GPU): ncu --set full of the k = 4 condensation kernel with the shared-memory factor
# (k_condense_b<4>, 10^6 random cells).  Call 25 imported the source into the report, which then exceeded the 64 MiB
# limit of gpurun_out/ and nothing came back; this version exports the raw and details pages and only keeps a small report.
mkdir -p gpurun_out
T=r2y
timeout 150 ncu --set full --clock-control none -k regex:k_condense_b -o gpurun_out/ncu_${T}_condense_b4 -f python tools/ncu_condense_k4.py 4 > gpurun_out/ncu_${T}.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/ncu_${T}.log
ncu -i gpurun_out/ncu_${T}_condense_b4.ncu-rep --page raw --csv > gpurun_out/ncu_${T}_condense_b4_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_${T}_condense_b4.ncu-rep --page details > gpurun_out/ncu_${T}_condense_b4_details.txt 2>/dev/null
ncu -i gpurun_out/ncu_${T}_condense_b4.ncu-rep --page source --csv > gpurun_out/ncu_${T}_condense_b4_source.csv 2>/dev/null
gzip -f gpurun_out/ncu_${T}_condense_b4_source.csv
sz=$(stat -c %s gpurun_out/ncu_${T}_condense_b4.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/ncu_${T}_condense_b4.ncu-rep; fi
ls -la gpurun_out/
echo done
