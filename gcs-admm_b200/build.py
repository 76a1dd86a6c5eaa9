"""Builds libgcsadmm.so (CUDA, sm_100a) in-tree.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "gcsadmm.cu")


def deps():
    """Every source the library is built from: a change to any of them triggers a rebuild in ``lib.load()``."""
    import glob
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh"))
                  + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))


OUT = os.path.join(HERE, "libgcsadmm.so")


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgcsadmm.so cannot be built")


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps()):
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgcsadmm.so")
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as fh:
        fh.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
