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


def deps():
    """every file the library is compiled from: all of csrc/ and all of include/ (globbed, so a new header cannot be
    forgotten — round 1 shipped a list that missed the headers holding the default kernels)"""
    import glob
    return sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh"))
                  + glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)])


def source_hash() -> str:
    """sha256 over the sources and the compiler flags; stored next to the .so so that `lib()` can refuse a binary that
    was built from other sources (copied trees carry arbitrary mtimes)."""
    import hashlib
    h = hashlib.sha256()
    for d in deps():
        if d.endswith("build.py"):
            continue
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS[:6]).encode())
    return h.hexdigest()


HASH_FILE = os.path.join(OUT_DIR, "libmbd_b200.sha256")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"),
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def built_hash() -> str:
    try:
        with open(HASH_FILE) as f:
            return f.read().strip()
    except OSError:
        return ""


def is_stale() -> bool:
    """True when the library is missing or was built from different sources (content hash, not mtimes)."""
    return not os.path.exists(OUT) or built_hash() != source_hash()


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
        with open(HASH_FILE + f".tmp{os.getpid()}", "w") as f:
            f.write(source_hash() + "\n")
        os.replace(HASH_FILE + f".tmp{os.getpid()}", HASH_FILE)
        if verbose:
            print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
