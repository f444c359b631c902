"""TEST INFRASTRUCTURE -- compiles the oracle's plain-C restatements (gcc + OpenMP) into oracle/_build/liboracle.so."""

from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SOURCES = ["terrain_oracle.c", "variogram_oracle.c"]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, *srcs, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed: {' '.join(cmd)}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
