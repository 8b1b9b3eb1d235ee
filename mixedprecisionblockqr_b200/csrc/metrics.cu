// metrics.cu — device-side parity metrics + the reference's CSV result log.
//
// The reference's harness metrics are O(m^3) single-thread host loops in FP32
// (h_strip_R_from_A / h_backward_error / h_q_error / h_lower_trapezoid_error, reference
// Cuda/qr.cu:85-196; h_matrix_norm, Cuda/mmult.cuh), unusable beyond ~2048^2.  Here the same
// quantities are computed on the GPU with FP64 accumulation:
//   * ||A - Q R||_F / ||A||_F and Q^T Q - I come from one tiled FP64 (DFMA) product kernel whose
//     epilogue reduces the tile on the fly, so neither Q R nor Q^T Q is ever written to HBM
//     (the reference allocates both, Cuda/qr.cu:117-118, :143-145);
//   * R is read through the mask row <= col, so the packed factor can be passed where the reference
//     passes the stripped R; Q^T Q is symmetric, only tiles on or above the diagonal are computed;
//   * per-CTA partials are folded by a one-CTA kernel in a fixed order: results are deterministic.
// Host part: h_qr_flops_per_second (Cuda/qr.cu:102-113) and h_write_results_to_log
// (Cuda/qr.cu:58-83, the "rows,cols,runtime,flops,error" file Cuda/performance/util.py:19-31 reads).
#include <errno.h>
#include <math.h>
#include <string.h>
#include <sys/stat.h>

#include <string>

#include "common.cuh"

namespace mpqr {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;  // CTA tile of the product, depth of one stage
constexpr int AS = TM + 2;                // row stride (doubles) of the A stage: keeps 16-byte alignment, spreads banks
constexpr int kThreads = 256;
constexpr int kNumPartial = 4;            // doubles per CTA: sum of squares, second sum, max signed, max abs

enum Mode { kBackward = 0, kOrtho = 1 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Folds (sum, sum, max, max) over the CTA and lets thread 0 store the four partials.
__device__ void cta_fold_store(double s0, double s1, double mx, double mxa, double* out) {
    __shared__ double red[kNumPartial][kThreads / 32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    mx = warp_max(mx);
    mxa = warp_max(mxa);
    if (l == 0) { red[0][w] = s0; red[1][w] = s1; red[2][w] = mx; red[3][w] = mxa; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) {
            red[0][0] += red[0][i];
            red[1][0] += red[1][i];
            red[2][0] = fmax(red[2][0], red[2][i]);
            red[3][0] = fmax(red[3][0], red[3][i]);
        }
        out[0] = red[0][0]; out[1] = red[1][0]; out[2] = red[2][0]; out[3] = red[3][0];
    }
}

// kBackward: P = Q R (Q m x m, R m x n masked to row <= col), epilogue d = A0 - P:
//            partial = (sum d^2, sum A0^2, -, max |d|).
// kOrtho:    P = Q^T Q (m x m), tiles with tj >= ti only, epilogue e = P - I:
//            partial = (sum e^2 [off-diagonal tiles twice], 0, max e (signed: the reference's
//            definition, Cuda/qr.cu:150-157), max |e|).
template <int MODE>
__global__ void __launch_bounds__(kThreads) product_metric_kernel(const float* __restrict__ Q, long ldq,
                                                                  const float* __restrict__ R, long ldr,
                                                                  const float* __restrict__ A0, long lda0, int m, int n,
                                                                  double* __restrict__ partial) {
    __shared__ __align__(16) double As[TK][AS];
    __shared__ __align__(16) double Bs[TK][TN];
    const int ti = blockIdx.y, tj = blockIdx.x;
    double* out = partial + ((size_t)ti * gridDim.x + tj) * kNumPartial;
    if (MODE == kOrtho && tj < ti) {
        if (threadIdx.x == 0) { out[0] = 0; out[1] = 0; out[2] = -INFINITY; out[3] = 0; }
        return;
    }
    const int i0 = ti * TM, j0 = tj * TN;
    const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    // depth of the product: Q R only needs k <= col (R is upper triangular), Q^T Q all m rows
    const int K = MODE == kBackward ? min(m, j0 + TN) : m;

    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = t + kThreads * q;
            if (MODE == kBackward) {  // A stage element (ii, kk) = Q[i0+ii, k0+kk]
                const int ii = e >> 4, kk = e & 15;
                ra[q] = (i0 + ii < m && k0 + kk < K) ? Q[(size_t)(i0 + ii) * ldq + k0 + kk] : 0.f;
            } else {  // (kk, ii) = Q[k0+kk, i0+ii]
                const int kk = e >> 6, ii = e & 63;
                ra[q] = (k0 + kk < K && i0 + ii < m) ? Q[(size_t)(k0 + kk) * ldq + i0 + ii] : 0.f;
            }
            const int kk = e >> 6, jj = e & 63;
            if (MODE == kBackward)
                rb[q] = (k0 + kk < K && j0 + jj < n && k0 + kk <= j0 + jj) ? R[(size_t)(k0 + kk) * ldr + j0 + jj] : 0.f;
            else
                rb[q] = (k0 + kk < K && j0 + jj < m) ? Q[(size_t)(k0 + kk) * ldq + j0 + jj] : 0.f;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = t + kThreads * q;
            if (MODE == kBackward) As[e & 15][e >> 4] = (double)ra[q];
            else As[e >> 6][e & 63] = (double)ra[q];
            Bs[e >> 6][e & 63] = (double)rb[q];
        }
    };

    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    fetch(0);
    for (int k0 = 0; k0 < K; k0 += TK) {
        stage();
        __syncthreads();
        if (k0 + TK < K) fetch(k0 + TK);  // next stage's global loads fly during the DFMAs
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const double2 a01 = *reinterpret_cast<const double2*>(&As[kk][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2*>(&As[kk][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4]);
            const double2 b23 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4 + 2]);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y};
            const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }

    double s0 = 0, s1 = 0, mx = -INFINITY, mxa = 0;
    const int rows = m, cols = MODE == kBackward ? n : m;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int i = i0 + ty * 4 + x;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int j = j0 + tx * 4 + y;
            if (i >= rows || j >= cols) continue;
            if (MODE == kBackward) {
                const double a0 = (double)A0[(size_t)i * lda0 + j];
                const double d = a0 - acc[x][y];
                s0 += d * d;
                s1 += a0 * a0;
                mxa = fmax(mxa, fabs(d));
            } else {
                const double e = acc[x][y] - (i == j ? 1.0 : 0.0);
                s0 += (tj > ti ? 2.0 : 1.0) * e * e;
                mx = fmax(mx, e);
                mxa = fmax(mxa, fabs(e));
            }
        }
    }
    cta_fold_store(s0, s1, mx, mxa, out);
}

// Elementwise passes over an m x n matrix; WHAT selects the quantity:
//   0: Frobenius: (sum x^2, 0, max x, max |x|)
//   1: strict lower trapezoid (col < row), h_lower_trapezoid_error (Cuda/qr.cu:173-196)
//   2: | |X| - |Y| | on row <= col: (sum d^2, sum Y^2, max d, max |Y|)   (north_star: elementwise |R| agreement)
template <int WHAT>
__global__ void __launch_bounds__(kThreads) elementwise_metric_kernel(const float* __restrict__ X, long ldx,
                                                                      const float* __restrict__ Y, long ldy, long m, long n,
                                                                      double* __restrict__ partial) {
    double s0 = 0, s1 = 0, mx = -INFINITY, mxa = 0;
    const long total = m * n;
    for (long e = (long)blockIdx.x * kThreads + threadIdx.x; e < total; e += (long)gridDim.x * kThreads) {
        const long i = e / n, j = e - i * n;
        if (WHAT == 0) {
            const double x = X[i * ldx + j];
            s0 += x * x;
            mx = fmax(mx, x);
            mxa = fmax(mxa, fabs(x));
        } else if (WHAT == 1) {
            if (j < i) {
                const double x = X[i * ldx + j];
                s0 += x * x;
                mxa = fmax(mxa, fabs(x));
            }
        } else {
            if (i <= j) {
                const double x = fabs((double)X[i * ldx + j]), y = fabs((double)Y[i * ldy + j]);
                const double d = fabs(x - y);
                s0 += d * d;
                s1 += y * y;
                mx = fmax(mx, d);
                mxa = fmax(mxa, y);
            }
        }
    }
    cta_fold_store(s0, s1, mx, mxa, partial + (size_t)blockIdx.x * kNumPartial);
}

__global__ void __launch_bounds__(kThreads) fold_partials_kernel(const double* __restrict__ partial, long count,
                                                                 double* __restrict__ out) {
    double s0 = 0, s1 = 0, mx = -INFINITY, mxa = 0;
    for (long i = threadIdx.x; i < count; i += kThreads) {
        s0 += partial[i * kNumPartial + 0];
        s1 += partial[i * kNumPartial + 1];
        mx = fmax(mx, partial[i * kNumPartial + 2]);
        mxa = fmax(mxa, partial[i * kNumPartial + 3]);
    }
    cta_fold_store(s0, s1, mx, mxa, out);
}

__global__ void strip_r_kernel(const float* __restrict__ A, long lda, float* __restrict__ R, long ldr, long m, long n) {
    const long total = m * n;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long i = e / n, j = e - i * n;
        R[i * ldr + j] = i <= j ? A[i * lda + j] : 0.f;
    }
}

struct Partials {
    double* d = nullptr;
    ~Partials() { if (d) cudaFree(d); }
    int alloc(long ctas) {
        if (cudaMalloc(&d, (size_t)(ctas + 1) * kNumPartial * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            set_error("metrics: device allocation of %ld partials failed", ctas);
            return MPQR_ENOMEM;
        }
        return MPQR_OK;
    }
};

// Folds `ctas` partials, copies the four results to host and waits for them.
int finish(Partials& p, long ctas, cudaStream_t st, double res[kNumPartial]) {
    double* out = p.d + (size_t)ctas * kNumPartial;
    fold_partials_kernel<<<1, kThreads, 0, st>>>(p.d, ctas, out);
    MPQR_CUDA(cudaGetLastError());
    MPQR_CUDA(cudaMemcpyAsync(res, out, kNumPartial * sizeof(double), cudaMemcpyDeviceToHost, st));
    MPQR_CUDA(cudaStreamSynchronize(st));
    return MPQR_OK;
}

int elementwise_grid(long total) {
    DeviceInfo di;
    if (get_device_info(&di) != MPQR_OK) return -1;
    long want = ceil_divl(total, (long)kThreads * 4);
    long cap = (long)di.num_sms * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace
}  // namespace mpqr

using namespace mpqr;

extern "C" int mpqr_strip_r_device(const float* dA_packed, long lda, float* dR, long ldr, int m, int n, void* stream) {
    if (!dA_packed || !dR || m < 1 || n < 1 || lda < n || ldr < n) {
        set_error("mpqr_strip_r_device: bad arguments");
        return MPQR_EINVAL;
    }
    const int grid = elementwise_grid((long)m * n);
    if (grid < 0) return MPQR_ECUDA;
    strip_r_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(dA_packed, lda, dR, ldr, m, n);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

extern "C" int mpqr_backward_error_device(const float* dA0, long lda0, const float* dR, long ldr, const float* dQ,
                                          long ldq, int m, int n, double* err, double* a_norm, void* stream) {
    if (!dA0 || !dR || !dQ || m < 1 || n < 1 || lda0 < n || ldr < n || ldq < m || !err) {
        set_error("mpqr_backward_error_device: bad arguments");
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(ceil_div(n, TN), ceil_div(m, TM));
    const long ctas = (long)grid.x * grid.y;
    Partials p;
    MPQR_TRY(p.alloc(ctas));
    product_metric_kernel<kBackward><<<grid, kThreads, 0, st>>>(dQ, ldq, dR, ldr, dA0, lda0, m, n, p.d);
    MPQR_CUDA(cudaGetLastError());
    double res[kNumPartial];
    MPQR_TRY(finish(p, ctas, st, res));
    const double an = sqrt(res[1]);
    if (a_norm) *a_norm = an;
    *err = an > 0 ? sqrt(res[0]) / an : sqrt(res[0]);
    return MPQR_OK;
}

extern "C" int mpqr_q_error_device(const float* dQ, long ldq, int m, double* max_signed, double* max_abs, double* fro,
                                   void* stream) {
    if (!dQ || m < 1 || ldq < m) {
        set_error("mpqr_q_error_device: bad arguments");
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(ceil_div(m, TN), ceil_div(m, TM));
    const long ctas = (long)grid.x * grid.y;
    Partials p;
    MPQR_TRY(p.alloc(ctas));
    product_metric_kernel<kOrtho><<<grid, kThreads, 0, st>>>(dQ, ldq, nullptr, 0, nullptr, 0, m, m, p.d);
    MPQR_CUDA(cudaGetLastError());
    double res[kNumPartial];
    MPQR_TRY(finish(p, ctas, st, res));
    if (fro) *fro = sqrt(res[0]);
    if (max_signed) *max_signed = res[2] > 0 ? res[2] : 0;  // the reference's running maximum starts at 0 (Cuda/qr.cu:150)
    if (max_abs) *max_abs = res[3];
    return MPQR_OK;
}

namespace {
template <int WHAT>
int elementwise_metric(const float* X, long ldx, const float* Y, long ldy, long m, long n, cudaStream_t st,
                       double res[kNumPartial]) {
    const int grid = elementwise_grid(m * n);
    if (grid < 0) return MPQR_ECUDA;
    Partials p;
    MPQR_TRY(p.alloc(grid));
    elementwise_metric_kernel<WHAT><<<grid, kThreads, 0, st>>>(X, ldx, Y, ldy, m, n, p.d);
    MPQR_CUDA(cudaGetLastError());
    return finish(p, grid, st, res);
}
}  // namespace

extern "C" int mpqr_lower_trapezoid_error_device(const float* dR, long ldr, int m, int n, double* err, void* stream) {
    if (!dR || m < 1 || n < 1 || ldr < n || !err) {
        set_error("mpqr_lower_trapezoid_error_device: bad arguments");
        return MPQR_EINVAL;
    }
    double res[kNumPartial];
    MPQR_TRY(elementwise_metric<1>(dR, ldr, nullptr, 0, m, n, (cudaStream_t)stream, res));
    *err = sqrt(res[0]);
    return MPQR_OK;
}

extern "C" int mpqr_frobenius_norm_device(const float* dX, long ldx, long rows, long cols, double* nrm, void* stream) {
    if (!dX || rows < 1 || cols < 1 || ldx < cols || !nrm) {
        set_error("mpqr_frobenius_norm_device: bad arguments");
        return MPQR_EINVAL;
    }
    double res[kNumPartial];
    MPQR_TRY(elementwise_metric<0>(dX, ldx, nullptr, 0, rows, cols, (cudaStream_t)stream, res));
    *nrm = sqrt(res[0]);
    return MPQR_OK;
}

extern "C" int mpqr_r_agreement_device(const float* dR, long ldr, const float* dRref, long ldref, int m, int n,
                                       double* max_abs_diff, double* max_abs_ref, double* fro_diff, void* stream) {
    if (!dR || !dRref || m < 1 || n < 1 || ldr < n || ldref < n) {
        set_error("mpqr_r_agreement_device: bad arguments");
        return MPQR_EINVAL;
    }
    double res[kNumPartial];
    MPQR_TRY(elementwise_metric<2>(dR, ldr, dRref, ldref, m, n, (cudaStream_t)stream, res));
    if (max_abs_diff) *max_abs_diff = res[2] < 0 ? 0 : res[2];
    if (max_abs_ref) *max_abs_ref = res[3];
    if (fro_diff) *fro_diff = sqrt(res[0]);
    return MPQR_OK;
}

// ------------------------------------------------------------------------------ host side
extern "C" float mpqr_qr_flops_per_second(float time_ms, int m, int n) {
    // the reference's own operation count 4 m^2 n - m n^2 + n^3 / 3 (Cuda/qr.cu:102-113), kept so that
    // log files stay comparable with the reference's plots; bench.py reports the Householder count.
    const float mf = (float)m, nf = (float)n;
    float flops = 4.0f * mf * mf * nf;
    flops -= mf * nf * nf;
    flops += nf * nf * nf / 3.0f;
    return flops / (time_ms / 1000.0f);
}

extern "C" int mpqr_write_results_to_log(const char* log_dir, const char* file_name, int height, int width, float time_ms,
                                         float flops_per_second, float backward_error) {
    // Same file format as h_write_results_to_log (Cuda/qr.cu:58-83): header once, then one line of
    // std::to_string(double) values, appended.  The reference hard-codes "log/<name>.txt".
    const std::string dir = (log_dir && *log_dir) ? log_dir : "log";
    const std::string name = (file_name && *file_name) ? file_name : "logFile";
    if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) {
        set_error("mpqr_write_results_to_log: cannot create %s: %s", dir.c_str(), strerror(errno));
        return MPQR_EINVAL;
    }
    const std::string path = dir + "/" + name + ".txt";
    struct stat sb;
    const bool fresh = stat(path.c_str(), &sb) != 0;
    FILE* f = fopen(path.c_str(), "a");
    if (!f) {
        set_error("mpqr_write_results_to_log: cannot open %s: %s", path.c_str(), strerror(errno));
        return MPQR_EINVAL;
    }
    std::string line;
    if (fresh) line += "rows,cols,runtime,flops,error\n";
    const double params[5] = {height * 1.0, width * 1.0, time_ms, flops_per_second, backward_error};
    for (int i = 0; i < 5; ++i) {
        line += std::to_string(params[i]);
        if (i != 4) line += ',';
    }
    line += "\n";
    const bool ok = fputs(line.c_str(), f) >= 0;
    if (fclose(f) != 0 || !ok) {
        set_error("mpqr_write_results_to_log: write to %s failed", path.c_str());
        return MPQR_EINVAL;
    }
    return MPQR_OK;
}
