"""Builds snesimage_b200/libsnesgpu.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

-fmad=false is part of the numerics contract (csrc/common.cuh): fused multiply-adds exist only where
the kernels spell them out, which is what keeps every f32 plane bit-identical to the CPU oracle.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsnesgpu.so")
SOURCES = ["snesgpu.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "lab.cuh", "dither.cuh", "dither_core.h", "kmeans.cuh", "score_common.cuh", "score_v2.cuh", "score_v3.cuh", "assign_delta.cuh", os.path.join("..", "..", "include", "snesgpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsnesgpu.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def scorer_source_hash() -> str:
    """sha256 over the sources that define k_score_v3: profiles/roofline_traffic.json records it, bench.py reports the
    file's DRAM traffic only while it still describes the kernel in the tree."""
    import hashlib
    h = hashlib.sha256()
    for f in ("common.cuh", "score_common.cuh", "score_v2.cuh", "score_v3.cuh"):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", SO, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.check_call(cmd)
    return SO
