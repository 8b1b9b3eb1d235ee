// ref_shim.cpp — the reference's own driver symbols on top of the C-ABI (libmpqr_refshim.so).
//
// The reference has no FFI layer: its operator API is C++ free functions with C-compatible
// signatures, declared in Cuda/qr.cuh:129-137 and called by test_dev_mixed_precision_block_qr
// (Cuda/qr.cu:1879) / test_dev_block_qr (Cuda/qr.cu:1826).  This translation unit defines the
// SAME (C++-mangled) symbols, so the reference's Cuda/main.cu + test harness can be linked
// against libmpqr_refshim.so instead of its own qr.cu drivers, unchanged.
//
// Contract kept: caller owns A ((m+1)*n floats, rows 0..m-1 = input, row m = 0) and
// Q (m*m floats); both are overwritten in place; errors print a message and exit(EXIT_FAILURE) like
// checkCudaErrors (Cuda/helper_cuda.h:583-595).  One deviation: the reference frees everything per call
// (Cuda/qr.cu:1221-1226), mpqr_block_qr_host keeps the plan of the last shape (handle + device copies,
// ~12 GB at 32768^2) for the next call; mpqr_release_cache() or MPQR_NO_HOST_CACHE=1 give the memory back.
#include <cstdio>
#include <cstdlib>

#include "../../include/mpqr.h"

namespace {
void run(float* A, float* Q, int m, int n, int r, unsigned flags, const char* who) {
    const int rc = mpqr_block_qr_host(A, Q, m, n, r, flags);
    if (rc != MPQR_OK) {
        std::fprintf(stderr, "CUDA error at %s code=%d \"%s\"\n", who, rc, mpqr_last_error());
        std::exit(EXIT_FAILURE);
    }
}
}  // namespace

// Cuda/qr.cu:1049-1226 (decl Cuda/qr.cuh:133): FP16 tensor-core path
void dev_mixed_precision_block_qr(float* A, float* Q, int m, int n, int r) { run(A, Q, m, n, r, MPQR_FP16, "dev_mixed_precision_block_qr"); }
// Cuda/qr.cu:958-1047: FP32 sibling with device WY
void dev_block_qr_wy(float* A, float* Q, int m, int n, int r) { run(A, Q, m, n, r, MPQR_FP32, "dev_block_qr_wy"); }
// Cuda/qr.cu:877-956: older FP32 variant, same contract
void dev_block_qr(float* A, float* Q, int m, int n, int r) { run(A, Q, m, n, r, MPQR_FP32, "dev_block_qr"); }
