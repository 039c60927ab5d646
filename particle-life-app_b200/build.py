"""Builds libplife.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python particle-life-app_b200/build.py [--force] [--verbose]

The .so lands in particle-life-app_b200/plife/libplife.so so that it travels
to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "plife", "libplife.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall", "-I", INCLUDE]
# per-file extra flags
EXTRA = {
    # the fp64 path must round every operation separately (reference JOML arithmetic is non-fused)
    "force_f64.cu": ["-fmad=false"],
    "slab.cu": [],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _deps_mtime():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [cc, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stdout + r.stderr, flush=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [cc, *ARCH, "-shared", "-o", OUT, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


def build_c_client() -> str:
    """gcc examples/plife_headless.c against include/plife.h and libplife.so: a plain C user of the ABI
    (no CUDA headers, no Python).  The binary is git-ignored and travels to the GPU box like the library."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "examples", "plife_headless.c")
    out = os.path.join(root, "examples", "plife_headless")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(OUT), os.path.getmtime(os.path.join(INCLUDE, "plife.h"))):
        return out
    libdir = os.path.dirname(OUT)
    cmd = ["gcc", "-O2", "-std=c11", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Wextra", "-Werror", "-I", INCLUDE, src, "-o", out,
           "-L", libdir, "-lplife", "-Wl,-rpath,$ORIGIN/../particle-life-app_b200/plife"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed on the C client:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
