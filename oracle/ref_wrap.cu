// oracle/ref_wrap.cu — TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" wrapper that is compiled TOGETHER with the UNMODIFIED reference
// sources (/root/reference/Cuda/qr.cu, mmult.cu — read where they lie, never copied)
// into oracle/_ref/libref_qr.so by oracle/build_ref.sh.  It only forwards to the
// reference's own C++ functions (declared in Cuda/qr.cuh:68-137) so that Python tests
// and bench.py's cpu_baseline / --impl reference legs can call them through ctypes.
//
// Nothing in the product path (mixedprecisionblockqr_b200/) links or loads this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#include "mmult.cuh"
#include "qr.cuh"

namespace {
// The reference prints from host metrics and (per thread!) from a device kernel
// (Cuda/qr.cu:500).  Silence stdout around reference calls.
struct StdoutMute {
    int saved;
    StdoutMute() {
        fflush(stdout);
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        dup2(devnull, 1);
        close(devnull);
    }
    ~StdoutMute() {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};
}  // namespace

extern "C" {

// Cuda/qr.cu:198  panel factorisation (host)
void ref_h_householder_qr(float* A, int m, int n, int global_offset, int panel_width) {
    h_householder_qr(A, m, n, global_offset, panel_width);
}

// Cuda/qr.cu:337  host WY; writes the dense (m-off)x(m-off) I - W Y^T into out
void ref_h_wy_transform(float* A, float* out, int m, int n, int global_offset, int panel_width) {
    float* q = nullptr;
    h_wy_transform(A, &q, m, n, global_offset, panel_width);
    size_t d = (size_t)(m - global_offset);
    memcpy(out, q, d * d * sizeof(float));
    free(q);
}

// Cuda/qr.cu:296  explicit Q by backward accumulation (host)
void ref_h_q_backward_accumulation(float* A, float* Q, int m, int n) {
    float* q = nullptr;
    h_q_backward_accumulation(A, &q, m, n);
    memcpy(Q, q, (size_t)m * m * sizeof(float));
    free(q);
}

// Cuda/qr.cu:1275  CPU statement of the hot path
void ref_h_block_qr(float* A, float* Q, int m, int n, int r) { h_block_qr(A, Q, m, n, r); }

// Cuda/qr.cu:958 / :1049  GPU drivers (need a GPU; device printf muted)
void ref_dev_block_qr_wy(float* A, float* Q, int m, int n, int r) {
    StdoutMute mute;
    dev_block_qr_wy(A, Q, m, n, r);
}
void ref_dev_mixed_precision_block_qr(float* A, float* Q, int m, int n, int r) {
    StdoutMute mute;
    dev_mixed_precision_block_qr(A, Q, m, n, r);
}

// Cuda/qr.cu:85-196 metrics
void ref_h_strip_R_from_A(float* A, float* R, int m, int n) { h_strip_R_from_A(A, R, m, n); }
float ref_h_backward_error(float* A, float* R, float* Q, int m, int n, int bits) {
    StdoutMute mute;
    return h_backward_error(A, R, Q, m, n, bits);
}
float ref_h_q_error(float* Q, int m, int bits) {
    StdoutMute mute;
    return h_q_error(Q, m, bits);
}
float ref_h_lower_trapezoid_error(float* R, int m, int n, int bits) {
    StdoutMute mute;
    return h_lower_trapezoid_error(R, m, n, bits);
}
float ref_h_qr_flops_per_second(float time_ms, int m, int n) { return h_qr_flops_per_second(time_ms, m, n); }

}  // extern "C"
