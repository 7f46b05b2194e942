"""In-tree build of the sm_100a shared library (video_segment_b200/libvsb200.so).

nvcc cross-compiles without a GPU.  The library has no torch dependency: it is a plain
C-ABI shared object (include/vsb200.h) loaded with ctypes.  -fmad=false keeps every float
operation rounding like the reference's scalar code (parity of edge weights / merge gates).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvsb200.so")
SOURCES = ["preprocess.cu", "edges.cu", "sort.cu", "merge.cu", "results.cu", "region_hist.cu", "region_stage.cu", "shard.cu", "shape.cu", "capi_kernels.cu", "engine.cu", "pb_io.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    # host-side float code (constraint walk, shape moments, tube areas) must round like the reference's scalar code on
    # every host ISA: no a*b+c contraction (GCC's default is -ffp-contract=fast where the target has FMA)
    "-Xcompiler", "-ffp-contract=off",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp", ".inc"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "vsb200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
