"""Builds sprsolve_b200/lib/libsprsolve_b200.so IN-TREE with nvcc for sm_100a.

    python -m sprsolve_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false keeps the element-wise arithmetic identical to
the reference's (Rust never contracts a*b+c into an FMA); -lineinfo maps ncu's source page to
this code.  NCCL is resolved at run time (dlopen), only its header is needed here.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsprsolve_b200.so")

SOURCES = [
    "spmv.cu", "create.cu", "ingest.cu", "dist.cu", "vecops.cu", "ops.cu", "solver_common.cu",
    "bicgstab.cu", "minres.cu", "gs_solver.cu", "gs_wave.cu", "capi.cu",
]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Wno-deprecated-gpu-targets",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> list[str]:
    # the image's default CXX (/opt/gcc) lacks some runtime specs; prefer the distro compiler
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _deps_mtime() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h", ".cu")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


# Per-file ptxas options.  spmv.cu: at -O2/-O3 ptxas re-schedules the row loop of the SpMV kernel
# to minimise live ranges and sinks every x gather down to its first use, which serialises the
# 8 gathers of a batch (one L2 round trip each); -O1 keeps the PTX order (8 column loads, 8
# gathers, then the multiply-add chain).  Verified in SASS and on the B200 (profiles/).
PTXAS_FLAGS = {"spmv.cu": os.environ.get("SPB_SPMV_PTXAS", "-O1")}


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, *_host_cxx(), "-c", os.path.join(CSRC, src), "-o", obj]
    if PTXAS_FLAGS.get(src):
        cmd.insert(1, f"-Xptxas={PTXAS_FLAGS[src]}")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # plain host link (no relocatable device code): only sm_100a cubins end up in the library
    cmd = [cxx, "-shared", "-fPIC", "-o", LIB, *objs, f"-L{cuda_lib}", f"-Wl,-rpath,{cuda_lib}", "-lcudart", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
