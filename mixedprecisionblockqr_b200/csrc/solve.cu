// solve.cu — least-squares solve on top of the factorisation.
//
// The reference leaves this unimplemented: dev_QR_Solver in Cuda/QR/Solver/solver.cu:39-87 is
// pseudocode for GVL 5.3.2 (x = R^-1 Q^T b) and python/linear_least_sqare.py:5-22 is the NumPy
// demo.  Here:  b <- Q^T b  by the stored compact-WY factors, panel by panel
// (b[lam:] -= Y_p T_p^T (Y_p^T b[lam:]), Y read from the packed factor, T_p from the handle), then
// back substitution with the upper triangle of the packed factor.  FP32 throughout; bandwidth-
// bound (the packed factor is read twice), launched per panel / per 128-column block.
#include "internal.h"

using namespace mpqr;

namespace {

constexpr int MAXRHS = 8;
constexpr int TB = 128;  // back-substitution block

// Y_p[i][c] (i = row - lam) lives at packed[(row + 1) * lda + lam + c] for i >= c, else 0.
__device__ __forceinline__ float y_at(const float* __restrict__ P, long lda, int lam, int row, int c) {
    return (row - lam >= c) ? P[(size_t)(row + 1) * lda + lam + c] : 0.f;
}

// z[c][j] += sum_rows Y[row][c] * b[row][j]      (thread <-> reflector c, rows split over the grid)
__global__ void __launch_bounds__(128) qt_dot_kernel(const float* __restrict__ P, long lda, int m, int lam, int pw,
                                                      const float* __restrict__ Bm, long ldb, int nrhs, float* __restrict__ z,
                                                      int rows_per_cta) {
    const int c = threadIdx.x;
    const int r0 = lam + blockIdx.x * rows_per_cta;
    int r1 = r0 + rows_per_cta;
    if (r1 > m) r1 = m;
    float acc[MAXRHS];
#pragma unroll
    for (int j = 0; j < MAXRHS; ++j) acc[j] = 0.f;
    if (c < pw) {
        for (int row = r0; row < r1; ++row) {
            const float y = y_at(P, lda, lam, row, c);
#pragma unroll
            for (int j = 0; j < MAXRHS; ++j)
                if (j < nrhs) acc[j] = fmaf(y, __ldg(&Bm[(size_t)row * ldb + j]), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < MAXRHS; ++j)
            if (j < nrhs) atomicAdd(&z[c * MAXRHS + j], acc[j]);
    }
}

// z' = T^T z (per CTA, tiny), then b[row][j] -= sum_c Y[row][c] z'[c][j]   (one warp per row)
__global__ void __launch_bounds__(256) qt_update_kernel(const float* __restrict__ P, long lda, int m, int lam, int pw,
                                                         const float* __restrict__ T, int ldt, const float* __restrict__ z,
                                                         float* __restrict__ Bm, long ldb, int nrhs, int rows_per_cta) {
    __shared__ float zs[kPanelMaxWidth][MAXRHS];
    __shared__ float zp[kPanelMaxWidth][MAXRHS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < pw * MAXRHS; idx += 256) zs[idx / MAXRHS][idx % MAXRHS] = z[idx];
    __syncthreads();
    for (int idx = tid; idx < pw * MAXRHS; idx += 256) {
        const int t = idx / MAXRHS, j = idx % MAXRHS;
        float v = 0.f;
        for (int u = 0; u <= t; ++u) v = fmaf(T[(size_t)u * ldt + t], zs[u][j], v);  // (T^T z)[t] = sum_{u<=t} T[u][t] z[u]
        zp[t][j] = v;
    }
    __syncthreads();
    const int r0 = lam + blockIdx.x * rows_per_cta;
    int r1 = r0 + rows_per_cta;
    if (r1 > m) r1 = m;
    for (int row = r0 + warp; row < r1; row += 8) {
        float acc[MAXRHS];
#pragma unroll
        for (int j = 0; j < MAXRHS; ++j) acc[j] = 0.f;
        for (int c = lane; c < pw; c += 32) {
            const float y = y_at(P, lda, lam, row, c);
#pragma unroll
            for (int j = 0; j < MAXRHS; ++j) acc[j] = fmaf(y, zp[c][j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < MAXRHS; ++j)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        if (lane < nrhs) {
            float v = acc[0];
#pragma unroll
            for (int j = 1; j < MAXRHS; ++j) v = (lane == j) ? acc[j] : v;
            Bm[(size_t)row * ldb + lane] -= v;
        }
    }
}

// x[j0:j0+nb) = R[j0:j0+nb, j0:j0+nb)^-1 c[j0:j0+nb)   (column-oriented back substitution in shared memory)
__global__ void __launch_bounds__(TB) trsv_block_kernel(const float* __restrict__ P, long lda, int j0, int nbk, float* __restrict__ Bm,
                                                         long ldb, int nrhs) {
    extern __shared__ float sm[];
    float* Rs = sm;                 // nbk x (TB + 1)
    float* cs = sm + TB * (TB + 1);  // nbk x MAXRHS
    const int t = threadIdx.x;
    for (int idx = t; idx < nbk * nbk; idx += TB) {
        const int i = idx / nbk, k = idx % nbk;
        Rs[i * (TB + 1) + k] = (k >= i) ? P[(size_t)(j0 + i) * lda + j0 + k] : 0.f;
    }
    for (int idx = t; idx < nbk * MAXRHS; idx += TB) {
        const int i = idx / MAXRHS, j = idx % MAXRHS;
        cs[idx] = (j < nrhs) ? Bm[(size_t)(j0 + i) * ldb + j] : 0.f;
    }
    __syncthreads();
    for (int k = nbk - 1; k >= 0; --k) {
        const float d = Rs[k * (TB + 1) + k];
        const float inv = (d != 0.f) ? 1.f / d : 0.f;  // a skipped (zero) column leaves a zero pivot: minimum-norm style 0
        if (t < MAXRHS) cs[k * MAXRHS + t] *= inv;
        __syncthreads();
        if (t < k) {
            const float r = Rs[t * (TB + 1) + k];
#pragma unroll
            for (int j = 0; j < MAXRHS; ++j) cs[t * MAXRHS + j] = fmaf(-r, cs[k * MAXRHS + j], cs[t * MAXRHS + j]);
        }
        __syncthreads();
    }
    for (int idx = t; idx < nbk * MAXRHS; idx += TB) {
        const int i = idx / MAXRHS, j = idx % MAXRHS;
        if (j < nrhs) Bm[(size_t)(j0 + i) * ldb + j] = cs[idx];
    }
}

// c[0:j0) -= R[0:j0, j0:j0+nbk) x[j0:j0+nbk)   (one warp per row)
__global__ void __launch_bounds__(256) trsv_update_kernel(const float* __restrict__ P, long lda, int j0, int nbk, float* __restrict__ Bm,
                                                           long ldb, int nrhs) {
    __shared__ float xs[TB][MAXRHS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < nbk * MAXRHS; idx += 256) {
        const int i = idx / MAXRHS, j = idx % MAXRHS;
        xs[i][j] = (j < nrhs) ? Bm[(size_t)(j0 + i) * ldb + j] : 0.f;
    }
    __syncthreads();
    for (int row = blockIdx.x * 8 + warp; row < j0; row += gridDim.x * 8) {
        float acc[MAXRHS];
#pragma unroll
        for (int j = 0; j < MAXRHS; ++j) acc[j] = 0.f;
        for (int k = lane; k < nbk; k += 32) {
            const float r = P[(size_t)row * lda + j0 + k];
#pragma unroll
            for (int j = 0; j < MAXRHS; ++j) acc[j] = fmaf(r, xs[k][j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < MAXRHS; ++j)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        if (lane < nrhs) {
            float v = acc[0];
#pragma unroll
            for (int j = 1; j < MAXRHS; ++j) v = (lane == j) ? acc[j] : v;
            Bm[(size_t)row * ldb + lane] -= v;
        }
    }
}

}  // namespace

extern "C" int mpqr_solve_device(mpqr_handle* h, const float* dA_packed, long lda, float* dB, long ldb, int nrhs, void* stream) {
    if (!h || !dA_packed || !dB || nrhs < 1 || nrhs > MAXRHS || ldb < nrhs || lda < h->n) {
        set_error("mpqr_solve_device: bad arguments (1 <= nrhs <= %d)", MAXRHS);
        return MPQR_EINVAL;
    }
    if (!h->factored || h->mg) {
        set_error("mpqr_solve_device: needs a completed single-GPU mpqr_factor_device on this handle");
        return MPQR_ESTATE;
    }
    if (h->m < h->n) {
        set_error("mpqr_solve_device: needs m >= n");
        return MPQR_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    const int m = h->m, n = h->n, r = h->r;
    float* z = h->S32;  // pw x MAXRHS scratch
    h->launches = 0;
    // ---- b <- Q^T b = Q_last^T ... Q_1^T b
    for (int lam = 0, p = 0; lam < h->kmax; lam += r, ++p) {
        const int pw = (lam + r < h->kmax) ? r : h->kmax - lam;
        const int D = m - lam;
        int rows = ceil_div(D, 2 * di.num_sms);
        if (rows < 32) rows = 32;
        const int grid = ceil_div(D, rows);
        MPQR_CUDA(cudaMemsetAsync(z, 0, (size_t)kPanelMaxWidth * MAXRHS * sizeof(float), st));
        qt_dot_kernel<<<grid, 128, 0, st>>>(dA_packed, lda, m, lam, pw, dB, ldb, nrhs, z, rows);
        MPQR_CUDA(cudaGetLastError());
        qt_update_kernel<<<grid, 256, 0, st>>>(dA_packed, lda, m, lam, pw, h->T + (size_t)p * r * r, r, z, dB, ldb, nrhs, rows);
        MPQR_CUDA(cudaGetLastError());
        h->launches += 2;
    }
    // ---- back substitution R x = (Q^T b)[0:n]
    const size_t smem = ((size_t)TB * (TB + 1) + (size_t)TB * MAXRHS) * sizeof(float);
    MPQR_TRY(func_attr_once((const void*)trsv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int j1 = n; j1 > 0;) {
        const int j0 = (j1 - 1) / TB * TB;
        const int nbk = j1 - j0;
        trsv_block_kernel<<<1, TB, smem, st>>>(dA_packed, lda, j0, nbk, dB, ldb, nrhs);
        MPQR_CUDA(cudaGetLastError());
        h->launches += 1;
        if (j0 > 0) {
            int grid = ceil_div(j0, 8);
            if (grid > 4 * di.num_sms) grid = 4 * di.num_sms;
            trsv_update_kernel<<<grid, 256, 0, st>>>(dA_packed, lda, j0, nbk, dB, ldb, nrhs);
            MPQR_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        j1 = j0;
    }
    return MPQR_OK;
}
