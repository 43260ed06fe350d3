"""Builds hy2dl_b200/csrc/*.cu into the in-tree C-ABI library libhy2dl_b200.so for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libhy2dl_b200.so")
OBJ = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
] + os.environ.get("HY2DL_NVCC_EXTRA", "").split()      # e.g. -DHY2DL_US_STAMPS (timeline instrumentation of the unit-split kernels)
# bucket / routing / loss kernels are latency- and HBM-bound: keep mul and add unfused so that the
# step arithmetic rounds like the reference's separate ATen ops
NO_FMAD = {"conceptual.cu", "routing.cu", "loss.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    deps = [os.path.join(CSRC, f) for f in srcs + sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh"))] + [os.path.join(HERE, "..", "include", "hy2dl_b200.h")]
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(deps)
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_one(f):
        obj = os.path.join(OBJ, f[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(["-fmad=false"] if f in NO_FMAD else []), "-c", os.path.join(CSRC, f), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {f}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs, "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
