#!/bin/bash
# GPU call 28 of round 2 (1 GPU): the rebuilt library with the shared-memory-factor kernels as the k >= 3 defaults --
# Poisson parity tests of all degrees, the new variant tests, the configs[3] element test, smoke(), and the A/B line of
# the condensation microbenchmark at the engine's defaults
mkdir -p gpurun_out
T=r2A
timeout 150 python -m pytest tests/test_engine_poisson_gpu.py tests/test_zz_lsmem_gpu.py "tests/test_timesteppers_gpu.py::test_imex_k3_ssp2_config3_element" -m gpu -q -x > gpurun_out/pytest_poisson_${T}.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_poisson_${T}.log
tail -n 4 gpurun_out/pytest_poisson_${T}.log
timeout 60 python tools/condense_bench.py --nc 1000000 --k 4 3 --lsmem --reps 3 > gpurun_out/condense_bench_${T}.jsonl 2> gpurun_out/condense_bench_${T}.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/condense_bench_${T}.jsonl
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${T}.log 2>&1; tail -n 2 gpurun_out/smoke_${T}.log
echo done
