"""Build the CUDA engine in-tree: ``libhdg_b200.so`` next to this file (sm_100a only)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhdg_b200.so")
SOURCES = ["hdg_engine.cu"]
HEADERS = ["hdg_local.cuh", "hdg_krylov.cuh", "hdg_poisson.cuh", "hdg_poisson_s.cuh", "hdg_flow.cuh", "hdg_mg.cuh", "hdg_tent.cuh", "hdg_advblock.cuh", "hdg_comm.cuh", "hdg_tracer.cuh",
           "hdg_tables.inc",
           os.path.join("..", "..", "include", "hdg_b200.h")]
STAMP = LIB + ".flags"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine is CUDA-only and cannot be built without it")


def _nccl_include() -> list:
    """-I for <nccl.h> (csrc/hdg_comm.cuh needs the types only; the library itself is bound with dlopen at run time):
    the system header if there is one, else the copy that ships with torch's nvidia-nccl wheel"""
    if os.path.exists("/usr/include/nccl.h"):
        return []
    try:
        import nvidia.nccl

        inc = os.path.join(os.path.dirname(nvidia.nccl.__file__ or list(nvidia.nccl.__path__)[0]), "include")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return ["-I", inc]
    except Exception:
        pass
    return []


def _command() -> list:
    cmd = [_nvcc(), *NVCC_FLAGS, *_nccl_include(), *(os.path.join(CSRC, s) for s in SOURCES), "-o", LIB]
    # development shortcut: HDG_DEV_DEGREES="2" compiles only k=2 (the shipped build has k=1..4)
    dev = os.environ.get("HDG_DEV_DEGREES")
    if dev:
        mask = sum(1 << int(k) for k in dev.split(","))
        cmd.insert(1, f"-DHDG_DEGREES={mask}")
    return cmd


def _stamp() -> str:
    """what the library was built with: flags and source names, not the checkout's absolute path (the tree is
    copied to another directory on the GPU box, which must not make a shipped library look stale)"""
    return " ".join(os.path.basename(a) if os.sep in a else a for a in _command()[1:])


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    # a library built with other flags (e.g. a development subset of the degrees) is stale
    if not os.path.exists(STAMP) or open(STAMP).read() != _stamp():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """compile csrc/*.cu -> libhdg_b200.so; returns the library path"""
    if not force and not needs_build():
        return LIB
    import fcntl

    # one build at a time (two pytest sessions, or a test session next to a manual build, would otherwise link into
    # the same file); whoever waited re-checks whether the work has been done meanwhile
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        cmd = _command()
        stamp = _stamp()
        tmp = LIB + f".tmp{os.getpid()}"
        cmd[cmd.index("-o") + 1] = tmp  # link next to the target, then rename atomically
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr, file=sys.stderr)
        os.replace(tmp, LIB)
        with open(STAMP, "w") as fh:
            fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
