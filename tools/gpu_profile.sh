#!/bin/bash
# One `ncu --set full` capture of ONE kernel, summarised ON THE BOX so that only small CSVs travel back.
#
#   gpurun --timeout 240 -- bash tools/gpu_profile.sh <tag> <kernel-regex> <skip> -- <command ...>
#   e.g. bash tools/gpu_profile.sh r2a_tracer_adv k_tracer_adv_t 1 -- python tools/tracer_bench.py 512 2
#
# Lesson of round 1 (profiles/summary_r1.md): three .ncu-rep files with --import-source exceeded gpurun's
# 64 MiB return limit and the whole gpurun_out/ of that call was dropped.  So: one kernel per invocation,
# the report stays in /tmp on the box, and only `--page raw --csv`, `--page source --csv` (gzip) and the
# `--page details` text come back.  Run the same <command> without ncu first (B200_PROFILING.md): a number
# printed under ncu is never a bench value.
set -u
tag=$1; regex=$2; skip=$3; shift 3
[ "$1" == "--" ] && shift
mkdir -p gpurun_out
rep=/tmp/ncu_${tag}
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s "${skip}" -c 1 \
    -f -o "${rep}" "$@" > "gpurun_out/ncu_${tag}.log" 2>&1
echo "ncu rc=$?" >> "gpurun_out/ncu_${tag}.log"
if [ -f "${rep}.ncu-rep" ]; then
  ncu -i "${rep}.ncu-rep" --page raw --csv > "gpurun_out/ncu_${tag}_raw.csv" 2>/dev/null
  ncu -i "${rep}.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/ncu_${tag}_source.csv.gz"
  ncu -i "${rep}.ncu-rep" --page details > "gpurun_out/ncu_${tag}_details.txt" 2>/dev/null
  python tools/ncu_summary.py "gpurun_out/ncu_${tag}_raw.csv" > "gpurun_out/ncu_${tag}_summary.txt" 2>&1
  ls -la "${rep}.ncu-rep" >> "gpurun_out/ncu_${tag}.log"
  rm -f "${rep}.ncu-rep"
fi
du -sh gpurun_out >> "gpurun_out/ncu_${tag}.log"
echo done
