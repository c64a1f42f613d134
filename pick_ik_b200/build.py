"""Builds libpik_b200.so (the C-ABI library of include/pik.h) in-tree with nvcc for sm_100a.

``--fmad=false``: the arithmetic contract of csrc/pik_device.cuh (only the fma() calls written in the
source are fused) is what makes results bit-identical to the CPU oracle.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libpik_b200.so")
SOURCES = ["pik_kernels.cu", "pik_api.cu", "pik_comm.cu", "pik_urdf.cpp"]
HEADERS = ["pik_device.cuh", "pik_kernels.cuh", "pik_types.h", "pik_host_robot.h", "pik_internal.h", os.path.join("..", "..", "include", "pik.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "--fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib

    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(LIB_PATH + ".hash"):
        return True
    with open(LIB_PATH + ".hash") as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [
        os.path.join(CSRC, f) for f in SOURCES
    ] + ["-ldl"]  # NCCL itself is dlopen'ed at the first multi-GPU call (pik_comm.cu)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpik_b200.so")
    with open(LIB_PATH + ".hash", "w") as fh:
        fh.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
