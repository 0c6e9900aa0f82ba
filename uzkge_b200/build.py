"""Build uzkge_b200/lib/libuzkge_cuda.so (nvcc, sm_100a) with the Makefile in uzkge_b200/csrc.

    python -m uzkge_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libuzkge_cuda.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "uzkge_cuda.h")]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force: bool = False, verbose: bool = False) -> str:
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    if force or _stale():
        jobs = str(min(4, os.cpu_count() or 1))
        out = None if verbose else subprocess.DEVNULL
        subprocess.check_call(["make", "-C", CSRC, "-j", jobs], stdout=out)
    if not os.path.exists(LIB):
        raise RuntimeError("libuzkge_cuda.so was not produced")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
