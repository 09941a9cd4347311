"""1-vs-N GPU parity through torchrun (one rank per GPU, NCCL): runs tests/dist/run_dist_check.py on
two GPUs when the box has them (SURVEY.md §4 tier T4).  Skipped on single-GPU boxes; the host-side
logic of the same path is covered on CPU by tests/test_partition.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_vs_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist", "run_dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    cases = [l for l in res.stdout.splitlines() if l.startswith('{"case"')]
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert len(cases) >= 10 and all('"ok": true' in c for c in cases), "\n".join(cases)
