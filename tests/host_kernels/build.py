"""g++ build of the host harnesses (test infrastructure; see cuda_shim.h)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "incompressibleeulerhdg_b200", "csrc")
FLAGS = ["-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-D__device__=", "-D__host__=",
         "-D__global__=", "-D__forceinline__=inline", "-D__launch_bounds__(...)=", "-I", CSRC, "-I", HERE]


def build(source: str, outdir: str) -> ctypes.CDLL:
    out = os.path.join(outdir, os.path.splitext(source)[0] + ".so")
    res = subprocess.run(["g++", *FLAGS, os.path.join(HERE, source), "-o", out], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr)
    return ctypes.CDLL(out)
