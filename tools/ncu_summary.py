#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: one line per profiled launch with the roofline metrics.
usage: python tools/ncu_summary.py gpurun_out/prof_*_raw.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum"]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        names, units = rows[hdr], rows[hdr + 1]
        for r in rows[hdr + 2:]:
            d = dict(zip(names, r))
            u = dict(zip(names, units))
            print(f"== {path}: {d.get('Kernel Name')}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:70s} {d[k]:>18s} {u[k]}")


if __name__ == "__main__":
    main()
