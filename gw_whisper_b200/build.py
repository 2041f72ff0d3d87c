"""Builds the C-ABI shared library (hand-written sm_100a CUDA) in-tree with nvcc.

`python -m gw_whisper_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles for sm_100a on a
machine without a GPU; the resulting libgww_b200.so travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgww_b200.so")
SOURCES = ["gww_api.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "attention_tc.cuh", "attention_persist.cuh", "elementwise.cuh", "logmel.cuh", "qfront.cuh",
           os.path.join("..", "..", "include", "gww.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC", "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
