"""Build the C oracle (oracle/oracle.c -> oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

There is no oracle/_ref: the reference is pure Rust on un-vendored arkworks forks and this image has no
cargo/rustc, so the reference itself is unbuildable here.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.c")
OUT = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wextra", "-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
