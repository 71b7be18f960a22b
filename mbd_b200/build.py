"""Builds libmbd_b200.so (sm_100a only) in-tree with nvcc.  No torch dependency in the .so.

  python -m mbd_b200.build          # build if stale
  python -m mbd_b200.build --force

-fmad=false: FMAs only where fmaf() is written (bit-exact parity contract, include/mbd_fp32.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "mbd_b200.cu")
OUT_DIR = os.path.join(PKG, "_C")
OUT = os.path.join(OUT_DIR, "libmbd_b200.so")
DEPS = [SRC, os.path.join(PKG, "csrc", "xpbd_device.cuh")] + [
    os.path.join(ROOT, "include", h) for h in ("mbd_b200.h", "mbd_fp32.h", "mbd_model.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"),
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)           # one builder at a time (ranks of a torchrun job)
        if not force and not is_stale():
            return OUT
        tmp = OUT + f".tmp{os.getpid()}"
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, OUT)                       # atomic: a concurrent reader never sees a half-written library
        if verbose:
            print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
