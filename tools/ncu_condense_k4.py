#!/usr/bin/env python
"""One launch of the k = 4 condensation kernels on 10^6 random cells for an ncu capture (register kernel, then the
shared-memory-factor kernel k_condense_b; tools/gpu_profile.sh style: run under `ncu -k regex:k_condense`)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from incompressibleeulerhdg_b200.engine import HDGEngine  # noqa: E402
from incompressibleeulerhdg_b200.mesh import RandomAffineCells  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eng = HDGEngine(RandomAffineCells(1_000_000), k)
for mask in (0, 1):
    eng.set_tuning("poisson_lsmem", mask)
    eng.setup_poisson(keep_local=True)
    eng.synchronize()
