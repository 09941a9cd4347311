#!/usr/bin/env python
"""Registers / stack / spill bytes per kernel from the log of `python -m incompressibleeulerhdg_b200.build --force -v`.

usage: python -m incompressibleeulerhdg_b200.build --force -v 2> build.log; python tools/ptxas_table.py build.log [old_table]
With a second argument the rows are compared with an earlier table (profiles/ptxas_rNN.txt) and the differences printed.
"""
import re
import subprocess
import sys


def table(log_path):
    log = open(log_path).read()
    ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n"
                      r"\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                      r"ptxas info\s+: Used (\d+) registers", log)
    names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
    rows = []
    for e, d in zip(ents, names):
        name = re.sub(r"^void ", "", d).split("(")[0]
        rows.append(f"{name:58s}{int(e[4]):6d}{int(e[1]):7d}{int(e[2]):10d}{int(e[3]):10d}")
    return sorted(rows)


def main():
    rows = table(sys.argv[1])
    print(f"{'kernel':58s}{'regs':>6s}{'stack':>7s}{'spill st':>10s}{'spill ld':>10s}")
    print("\n".join(rows))
    if len(sys.argv) > 2:
        old = sorted(l.rstrip("\n") for l in open(sys.argv[2]) if not l.startswith("#") and not l.startswith("kernel"))
        gone, new = sorted(set(old) - set(rows)), sorted(set(rows) - set(old))
        print(f"\n# {len(rows)} kernels; {len(gone)} rows of {sys.argv[2]} changed or disappeared, {len(new)} new or changed rows",
              file=sys.stderr)
        for r in gone:
            print("# - " + r, file=sys.stderr)
        for r in new:
            print("# + " + r, file=sys.stderr)


if __name__ == "__main__":
    main()
