#!/bin/bash
# builds the development microbenchmarks next to their sources (binaries are git-ignored, travel with gpurun)
set -e
cd "$(dirname "$0")"
for src in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -I ../../incompressibleeulerhdg_b200/csrc -o "${src%.cu}.bin" "$src" -Xptxas -v 2> "${src%.cu}.ptxas.log"
done
