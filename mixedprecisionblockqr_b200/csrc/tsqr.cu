// tsqr.cu — tall-skinny QR — device restatement of ts_qr
// (reference python/ca_qr.py:25-43): split A (m x n, m >> n) into row blocks, factor every
// block independently, stack the n x n R factors, factor the stack, and (optionally) form the
// thin Q = blkdiag(Q_i[:, :n]) * Q_stack.  The reference fixes 4 blocks and a 2-level binary
// tree (ca_qr.py:27-38); here the block height is chosen so that a block's panels stay resident
// in shared memory (<= 32768 rows) and the stack is factored in one more level (recursively if
// it is itself tall).  Block factorisations use the FP32 driver (panel kernel + SIMT GEMMs):
// TSQR is bandwidth-bound (64 flop/B at n = 256), not tensor-bound.
//
// Round-1 status: functional and parity-tested; blocks are processed one after the other, so
// the GPU is latency-bound on the panel kernel.  A batched CTA-per-block kernel is the planned
// replacement ("next").
#include "internal.h"

using namespace mpqr;

namespace mpqr {
int sgemm_nn_store(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K,
                   cudaStream_t stream);
}

namespace {

__global__ void copy_block_kernel(const float* __restrict__ src, long lds, float* __restrict__ dst, long ldd, long rows, int cols) {
    long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long i = idx / cols;
        int j = (int)(idx - i * cols);
        dst[i * ldd + j] = src[i * lds + j];
    }
}

// dst (n x n, ldd) = upper triangle of the packed factor (rows <= cols), zero below
__global__ void extract_r_kernel(const float* __restrict__ packed, long ldp, float* __restrict__ dst, long ldd, int n) {
    int total = n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int i = idx / n, j = idx - i * n;
        dst[(size_t)i * ldd + j] = (i <= j) ? packed[(size_t)i * ldp + j] : 0.f;
    }
}

int grid_of(long total) {
    long g = (total + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

// thin Q (rows x n) of a factored block: backward accumulation on [I_n; 0]
int form_thin_q(mpqr_handle* h, float* Q, long ldq, cudaStream_t st) {
    const int m = h->m, n = h->n;
    MPQR_CUDA(cudaMemset2DAsync(Q, ldq * sizeof(float), 0, (size_t)n * sizeof(float), m, st));
    MPQR_TRY(set_identity(Q, ldq, n < m ? n : m, st));
    for (int p = h->npanels - 1; p >= 0; --p) {
        const int lam = p * h->r;
        const int pw = (lam + h->r < h->kmax) ? h->r : h->kmax - lam;
        const int D = m - lam, nc = n - lam;
        float* Y = h->Y32 + (size_t)lam * h->ld32 + lam;
        float* W = h->W32 + (size_t)lam * h->ld32 + lam;
        float* Qs = Q + (size_t)lam * ldq + lam;
        MPQR_TRY(sgemm_tn(Y, h->ld32, Qs, ldq, h->S32, h->lds32, pw, nc, D, st, &h->launches));
        MPQR_TRY(sgemm_nn_sub(W, h->ld32, h->S32, h->lds32, Qs, ldq, D, nc, pw, st, &h->launches));
    }
    return MPQR_OK;
}

struct Level {
    long rows;
    int nblk;
    long h;  // rows per block (last block may be shorter)
};

}  // namespace

extern "C" int mpqr_tsqr_device(const float* dA, long lda, long m, int n, float* dQ, long ldq, float* dR, long ldr,
                                void* stream) {
    if (!dA || !dR || n < 1 || m < n || lda < n || ldr < n || (dQ && ldq < n) || m > 0x7fffffffL) {
        set_error("mpqr_tsqr_device: bad arguments (needs m >= n)");
        return MPQR_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long HMAX = 32768;
    const int r = n < 128 ? n : 128;
    // single block: plain blocked QR
    long nblk = (m + HMAX - 1) / HMAX;  // block height <= HMAX: panels stay in the register-resident kernel
    if (nblk < 1 || m < 2L * n) nblk = 1;
    long hrows = (m + nblk - 1) / nblk;
    if (hrows < n) { nblk = 1; hrows = m; }
    while (nblk > 1 && m - (nblk - 1) * hrows < n) { --nblk; hrows = (m + nblk - 1) / nblk; }

    const long ldp = round_up(n, 4);
    float *work = nullptr, *rstack = nullptr, *qblk = nullptr, *qstack = nullptr;
    mpqr_handle *hb = nullptr, *hl = nullptr;
    int rc = MPQR_OK;
    const unsigned flags = MPQR_FP32 | (dQ ? MPQR_KEEP_WY : 0u);
    const long hlast = m - (nblk - 1) * hrows;
    do {
        if (cudaMalloc(&work, (size_t)(hrows + 1) * ldp * sizeof(float)) != cudaSuccess) { set_error("tsqr: alloc failed"); rc = MPQR_ENOMEM; break; }
        if (nblk > 1 && cudaMalloc(&rstack, (size_t)nblk * n * ldp * sizeof(float)) != cudaSuccess) { set_error("tsqr: alloc failed"); rc = MPQR_ENOMEM; break; }
        if ((rc = mpqr_create(&hb, (int)hrows, n, r, 0, flags))) break;
        if (hlast != hrows && (rc = mpqr_create(&hl, (int)hlast, n, r, 0, flags))) break;
        if (dQ && nblk > 1) {
            if (cudaMalloc(&qblk, (size_t)hrows * ldp * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&qstack, (size_t)nblk * n * ldp * sizeof(float)) != cudaSuccess) { set_error("tsqr: alloc failed"); rc = MPQR_ENOMEM; break; }
        }
        // Pass 1: R factor of every block (Q needs a second pass after the tree is known).
        for (long b = 0; b < nblk && rc == MPQR_OK; ++b) {
            const long rows = (b == nblk - 1) ? hlast : hrows;
            mpqr_handle* h = (rows == hrows) ? hb : hl;
            copy_block_kernel<<<grid_of(rows * n), 256, 0, st>>>(dA + (size_t)b * hrows * lda, lda, work, ldp, rows, n);
            if ((rc = mpqr_factor_device(h, work, ldp, st))) break;
            if (nblk == 1) {
                extract_r_kernel<<<grid_of((long)n * n), 256, 0, st>>>(work, ldp, dR, ldr, n);
                if (dQ) rc = form_thin_q(h, dQ, ldq, st);
            } else {
                extract_r_kernel<<<grid_of((long)n * n), 256, 0, st>>>(work, ldp, rstack + (size_t)b * n * ldp, ldp, n);
            }
        }
        if (rc != MPQR_OK || nblk == 1) break;
        // Tree: factor the stacked R's ((nblk*n) x n) — recursion handles a tall stack.
        if ((rc = mpqr_tsqr_device(rstack, ldp, nblk * (long)n, n, dQ ? qstack : nullptr, ldp, dR, ldr, st))) break;
        if (!dQ) break;
        // Pass 2: thin Q rows of block b = Q_b[:, :n] * Qstack[b*n:(b+1)*n, :]
        for (long b = 0; b < nblk && rc == MPQR_OK; ++b) {
            const long rows = (b == nblk - 1) ? hlast : hrows;
            mpqr_handle* h = (rows == hrows) ? hb : hl;
            copy_block_kernel<<<grid_of(rows * n), 256, 0, st>>>(dA + (size_t)b * hrows * lda, lda, work, ldp, rows, n);
            if ((rc = mpqr_factor_device(h, work, ldp, st))) break;
            if ((rc = form_thin_q(h, qblk, ldp, st))) break;
            rc = sgemm_nn_store(qblk, ldp, qstack + (size_t)b * n * ldp, ldp, dQ + (size_t)b * hrows * ldq, ldq, (int)rows, n, n, st);
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == MPQR_OK && e != cudaSuccess) { set_error("tsqr: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; }
    if (hb) mpqr_destroy(hb);
    if (hl) mpqr_destroy(hl);
    cudaFree(work); cudaFree(rstack); cudaFree(qblk); cudaFree(qstack);
    return rc;
}
