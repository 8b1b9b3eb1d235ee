"""Builds libmpqr.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo
snapshot to the GPU box.  `python -m mixedprecisionblockqr_b200.build [--force] [-v]`.
Every .cu is compiled to its own object (in parallel, only when stale) and linked."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(_HERE, "build")
LIB = os.path.join(_HERE, "libmpqr.so")
SHIM = os.path.join(_HERE, "libmpqr_refshim.so")
SOURCES = ["api.cu", "panel.cu", "panel_legacy.cu", "gemm_simt.cu", "gemm_tc.cu", "mg.cu", "tsqr.cu", "solve.cu", "loader.cu", "metrics.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "internal.h"), os.path.join(_HERE, "..", "include", "mpqr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs, jobs = [], []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            jobs.append([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for rc in ex.map(subprocess.call, jobs):
                if rc != 0:
                    raise RuntimeError("nvcc failed")
    if jobs or force or _stale(LIB, objs):
        subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB])
    shim_src = os.path.join(CSRC, "ref_shim.cpp")
    if os.path.exists(shim_src) and (force or _stale(SHIM, [shim_src, LIB])):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", shim_src, "-o", SHIM,
                               "-L" + _HERE, "-lmpqr", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
