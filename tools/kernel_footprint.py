#!/usr/bin/env python
"""Static footprint of every kernel in libhdg_b200.so: SASS instructions, code bytes (16 per instruction), registers,
stack (spills), static shared memory, and the FP64 / shared / local / global instruction mix.  Written after the ncu
capture of k_condense_b<4> (profiles/r2/ncu_r2y_condense_b4_*) showed an instruction-fetch bound: B200's instruction
caches hold 6 KB (L0, per scheduler) and 32 KB (L1.5, per SM) -- /opt/skills/guides/B300_MICROARCH.md "I-cache".

usage: python tools/kernel_footprint.py [path/to/libhdg_b200.so] > profiles/r2/kernel_footprint_r2.md
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "incompressibleeulerhdg_b200",
                                                         "libhdg_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
        name = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
mix = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        c = mix[cur]
        c["n"] += 1
        if op in ("DFMA", "DMUL", "DADD"):
            c["fp64"] += 1
        elif op in ("LDS", "STS"):
            c["shared"] += 1
        elif op in ("LDL", "STL"):
            c["local"] += 1
        elif op in ("LDG", "STG"):
            c["global"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(mix), capture_output=True, text=True).stdout.splitlines()
rows = []
for mangled, pretty in zip(mix, demangle):
    c = mix[mangled]
    reg, stack, shared = usage.get(mangled, (0, 0, 0))
    short = re.sub(r"\(.*", "", pretty).replace("void ", "")
    rows.append((c["n"], short, reg, stack, shared, c))
rows.sort(reverse=True)
print("# Static kernel footprint of libhdg_b200.so (tools/kernel_footprint.py)\n")
print("Instruction caches of a B200 SM: L0 6 KB per scheduler, L1.5 32 KB per SM (= 2 048 instructions of 16 bytes).  A kernel")
print("whose loop body is larger streams its code from L2 once per warp and pass; with few warps per SM that shows up as")
print("\"no instruction\" stalls (measured: k_condense_b<4>, 64 % of the stall cycles).\n")
print("| kernel | SASS instructions | code KB | registers | stack B | static smem B | FP64 | LDS/STS | LDL/STL | LDG/STG |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, short, reg, stack, shared, c in rows:
    if n < 1024:
        continue
    print(f"| `{short}` | {n} | {n * 16 / 1024:.0f} | {reg} | {stack} | {shared} | {c['fp64']} | {c['shared']} | {c['local']} | {c['global']} |")
small = [r for r in rows if r[0] < 1024]
print(f"\n{len(small)} further kernels have fewer than 1 024 instructions (16 KB).")
