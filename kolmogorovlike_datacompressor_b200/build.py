"""Build libkolm_b200.so in-tree with nvcc for sm_100a (one translation unit: csrc/abi.cu)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "abi.cu")
SO = os.path.join(HERE, "libkolm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(os.path.dirname(HERE), "include", "kolm_abi.h")]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    cmd = [NVCC] + FLAGS + ["-o", SO, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(HERE, "csrc", "ptxas_v.log"), "w") as f:
        f.write(r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
