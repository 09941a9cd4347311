#!/bin/bash
# 1-GPU job: full GPU test suite, default bench, launch list (ncu --metrics gpu__time_duration.sum)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
timeout 900 python bench.py > gpurun_out/bench_r1k_1gpu.json 2> gpurun_out/bench_r1k_1gpu.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1k_reference.json 2> gpurun_out/bench_r1k_reference.err
BENCH="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$BENCH > gpurun_out/plain_k.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file /tmp/launches_r1k.csv $BENCH > gpurun_out/ncu_k.log 2>&1
python tools/summarize_launches.py /tmp/launches_r1k.csv > gpurun_out/launches_r1k.md 2>&1
gzip -c /tmp/launches_r1k.csv > gpurun_out/launches_r1k.csv.gz
echo done
