"""Builds the product library in-tree: cerbos_b200/_lib/libcerbos_b200.so (sm_100a, -lineinfo).

    python -m cerbos_b200.csrc.build [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(ROOT, "cerbos_b200", "_lib")
SO = os.path.join(OUT_DIR, "libcerbos_b200.so")
SOURCES = [os.path.join(HERE, "cerbos_b200.cu")]
HEADERS = [os.path.join(HERE, "cb_core.h"), os.path.join(ROOT, "include", "cerbos_b200.h"),
           os.path.join(ROOT, "include", "cerbos_b200_format.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE,
           "-o", SO] + SOURCES
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose="--verbose" in sys.argv or "-v" in sys.argv))
