"""Build libzs.so in-tree with nvcc for sm_100a (no other architecture is compiled).

    python -m ossid_code_b200.build [--force]

The shared object lands next to this file so that it travels with the source tree to the
GPU box; a JIT cache elsewhere would not.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzs.so")
STAMP = os.path.join(HERE, "libzs.so.stamp")
SOURCES = ("zs_api.cu", "zs_features.cu", "zs_score_f32.cu", "zs_score_tc.cu", "zs_score_tc3.cu", "zs_head_tc.cu", "zs_topk.cu", "zs_metrics.cu", "zs_icp.cu")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    # keep IEEE division / sqrt and no flush-to-zero; exact ops use intrinsics (zs_common.cuh)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libzs.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "zs.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
