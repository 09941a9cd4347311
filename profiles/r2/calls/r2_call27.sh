#!/bin/bash
# GPU call 27 of round 2 (1 GPU): development microbenchmark of the k = 3, 4 condensation variants (tools/dev/condense_mb.cu:
# warps per block passing the phases together behind barriers x rows of a facet solved together), 10^6 random cells
mkdir -p gpurun_out
timeout 120 tools/dev/condense_mb.bin 1000000 3 > gpurun_out/condense_mb_r2z.jsonl 2> gpurun_out/condense_mb_r2z.err; echo "rc=$?"
cat gpurun_out/condense_mb_r2z.jsonl | cut -c1-260
tail -n 3 gpurun_out/condense_mb_r2z.err
