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
LIB_BF16 = os.path.join(HERE, "libgww_b200_bf16.so")
SOURCES = ["gww_api.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "attention_tc.cuh", "attention_persist.cuh", "elementwise.cuh", "logmel.cuh", "qfront.cuh", "whiten.cuh", "qadapter_tc.cuh",
           os.path.join("..", "..", "include", "gww.h")]


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f))


def build_lib(force: bool = False, verbose: bool = False, bf16: bool = False) -> str:
    """Builds libgww_b200.so (fp16 tensor-core operands, the default) or, with bf16=True,
    libgww_b200_bf16.so (bf16 operands; selected at run time with GWW_OPERAND=bf16)."""
    lib = LIB_BF16 if bf16 else LIB
    # tuning builds: extra -D switches into another file (selected at run time with GWW_LIB=<path>)
    extra = os.environ.get("GWW_BUILD_DEFS", "").split()
    if extra:
        lib = os.environ.get("GWW_BUILD_OUT") or os.path.join(HERE, "variants", "libgww_variant.so")
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        force = True
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC", "-o", lib,
    ] + (["-DGWW_OPERAND_BF16=1"] if bf16 else []) + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return lib


def build_all(force: bool = False, verbose: bool = False):
    """Both operand variants, compiled in parallel."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as ex:
        futs = [ex.submit(build_lib, force, verbose, b) for b in (False, True)]
        return [f.result() for f in futs]


if __name__ == "__main__":
    if "--bf16" in sys.argv:
        print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv, bf16=True))
    elif "--all" in sys.argv:
        print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
    else:
        print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
