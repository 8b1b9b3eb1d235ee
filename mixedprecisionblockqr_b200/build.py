"""Builds libmpqr.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo
snapshot to the GPU box.  `python -m mixedprecisionblockqr_b200.build [--force]`."""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libmpqr.so")
SHIM = os.path.join(_HERE, "libmpqr_refshim.so")
SOURCES = ["api.cu", "panel.cu", "panel_legacy.cu", "gemm_simt.cu", "gemm_tc.cu", "mg.cu", "tsqr.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not os.path.exists(os.path.join(CSRC, "mg.cu")):
        srcs.append(os.path.join(CSRC, "mg_stub.cu"))
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(_HERE, "..", "include", "mpqr.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB]
        subprocess.check_call(cmd)
    shim_src = os.path.join(CSRC, "ref_shim.cpp")
    if os.path.exists(shim_src) and (force or _stale(SHIM, [shim_src, LIB])):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", shim_src, "-o", SHIM,
                               "-L" + _HERE, "-lmpqr", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
