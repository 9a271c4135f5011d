"""Build the sm_100a shared library (C-ABI, include/pemap.h) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpemap.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-fmad=false", "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(ROOT, "include", "pemap.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "pemap_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed (see %s)" % log)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
