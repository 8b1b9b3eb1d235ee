// tsqr.cu — tall-skinny QR — device restatement of ts_qr
// (reference python/ca_qr.py:25-43): split A (m x n, m >> n) into row blocks, factor every
// block independently, stack the n x n R factors, factor the stack, and (optionally) form the
// thin Q = blkdiag(Q_i[:, :n]) * Q_stack.  The reference fixes 4 blocks and a 2-level binary
// tree (ca_qr.py:27-38); here the block height is chosen so that a block's panels stay resident
// in shared memory (<= 32768 rows) and the stack is factored in one more level (recursively if
// it is itself tall).  Block factorisations use the FP32 driver (panel kernel + SIMT GEMMs):
// TSQR is bandwidth-bound (64 flop/B at n = 256), not tensor-bound.
//
// The row blocks are independent: they are spread over up to 4 LANES (MPQR_TSQR_LANES) (stream + handle + buffers
// each, SM budget = device / lanes), so several blocks' register-block clusters (16 SMs each) run
// at the same time instead of one latency-bound panel chain.  Every block is factored ONCE and its
// reflectors (Y_b, W_b: FP32, m x n each for the whole matrix) stay resident until the tree is known;
// the thin Q rows of block b are then Q_b [Qstack_b; 0], i.e. the block's panels applied to its n x n
// slice of the stack's Q instead of to the identity: one pass, no separate Q_b and no m x n x n product.
#include <mutex>

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

// thin Q (rows x n) of a factored block: backward accumulation on [X; 0], X = I_n (seed == null) or the n x n
// matrix `seed` (the block's slice of the stack's Q: the rows come out as Q_b [seed; 0]).  Y / W: the block's
// reflectors (m x ld32, element (0, 0)).
int form_thin_q(mpqr_handle* h, const float* Yall, const float* Wall, float* Q, long ldq, const float* seed, long ldseed, cudaStream_t st) {
    const int m = h->m, n = h->n;
    MPQR_CUDA(cudaMemset2DAsync(Q, ldq * sizeof(float), 0, (size_t)n * sizeof(float), m, st));
    if (seed) MPQR_CUDA(cudaMemcpy2DAsync(Q, ldq * sizeof(float), seed, ldseed * sizeof(float), (size_t)n * sizeof(float), n < m ? n : m,
                                          cudaMemcpyDeviceToDevice, st));
    else MPQR_TRY(set_identity(Q, ldq, n < m ? n : m, st));
    for (int p = h->npanels - 1; p >= 0; --p) {
        const int lam = p * h->r;
        const int pw = (lam + h->r < h->kmax) ? h->r : h->kmax - lam;
        const int D = m - lam;
        // identity seed: columns left of lam are untouched by panel p; a dense seed fills all n columns
        const int cofs = seed ? 0 : lam, nc = n - cofs;
        // rows below n are still zero when the LAST panel is applied first: its product Y^T Q only needs rows lam .. n
        const int Dk = (p == h->npanels - 1 && n - lam < D) ? n - lam : D;
        const float* Y = Yall + (size_t)lam * h->ld32 + lam;
        const float* W = Wall + (size_t)lam * h->ld32 + lam;
        float* Qs = Q + (size_t)lam * ldq + cofs;
        MPQR_TRY(sgemm_tn(Y, h->ld32, Qs, ldq, h->S32, h->lds32, pw, nc, Dk, st, &h->launches));
        MPQR_TRY(sgemm_nn_sub(W, h->ld32, h->S32, h->lds32, Qs, ldq, D, nc, pw, st, &h->launches));
    }
    return MPQR_OK;
}

struct Lane {
    cudaStream_t s = nullptr;   // own stream (always: a plan outlives the caller's stream)
    cudaEvent_t done = nullptr;
    float* work = nullptr;   // (hrows + 1) x ldp packed factor of the current block
    mpqr_handle* hb = nullptr;
    mpqr_handle* hl = nullptr;  // handle of the (shorter) last block, on the lane that owns it
};

int tsqr_lanes(long nblk) {
    const char* e = getenv("MPQR_TSQR_LANES");
    long want = e ? atol(e) : 8;  // [B200, r2t] 1048576 x 256, stream-ordered lanes (see build_plan): 23.6 ms with 4 lanes, 21.7 ms with 8
    if (want < 1) want = 1;
    if (want > 16) want = 16;
    return (int)(want < nblk ? want : nblk);
}

// Everything a call needs besides its arguments: lanes (stream, handles, buffers), the R stack and the
// stack's Q.  Plans are cached per (device, m, n, with-Q, lanes): creating one costs ~10 cudaMalloc and a
// memset per lane, more than the factorisation of a small matrix.  mpqr_tsqr_release_cache() frees them.
struct Plan {
    int dev = -1;
    long m = 0;
    int n = 0, NL = 0;
    bool with_q = false;
    long nblk = 1, hrows = 0, hlast = 0, ldp = 0;
    int budget = 0;
    std::vector<Lane> lanes;
    float *rstack = nullptr, *qstack = nullptr;
    float *Yall = nullptr, *Wall = nullptr;   // nblk x hrows x ld32 each: every block's reflectors (with-Q plans, nblk > 1)
    long ld32 = 0;
    cudaEvent_t ev_start = nullptr, ev_tree = nullptr;
    bool busy = false;  // a recursive call never shares its caller's plan (different m), but be explicit
    ~Plan() {
        for (auto& L : lanes) {
            if (L.hb) mpqr_destroy(L.hb);
            if (L.hl) mpqr_destroy(L.hl);
            cudaFree(L.work);
            if (L.done) cudaEventDestroy(L.done);
            if (L.s) cudaStreamDestroy(L.s);
        }
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_tree) cudaEventDestroy(ev_tree);
        cudaFree(rstack); cudaFree(qstack); cudaFree(Yall); cudaFree(Wall);
    }
};
std::mutex g_plans_mu;
std::vector<Plan*> g_plans;
constexpr size_t kMaxPlans = 8;

int build_plan(Plan* P, const DeviceInfo& di) {
    constexpr long HMAX = 32768;   // leaf height: what one register-resident cluster holds
    const long m = P->m;
    const int n = P->n;
    const int r = n < 128 ? n : 128;
    long nblk = (m + HMAX - 1) / HMAX;  // block height <= HMAX: panels stay in the register-resident kernel
    if (nblk < 1 || m < 2L * n) nblk = 1;  // single block: plain blocked QR
    long hrows = (m + nblk - 1) / nblk;
    if (hrows < n) { nblk = 1; hrows = m; }
    while (nblk > 1 && m - (nblk - 1) * hrows < n) { --nblk; hrows = (m + nblk - 1) / nblk; }
    P->nblk = nblk; P->hrows = hrows; P->hlast = m - (nblk - 1) * hrows; P->ldp = round_up(n, 4);
    P->NL = tsqr_lanes(nblk);
    P->budget = P->NL > 1 ? (di.num_sms / P->NL < 16 ? 16 : di.num_sms / P->NL) : 0;
    P->lanes.resize(P->NL);
    // TSQR handles never use the persistent panel kernel (MPQR_STREAM_ORDERED): a persistent cluster waits on device flags for
    // side-stream work issued behind it, and with several lanes' clusters resident while this driver went on to build the plan of
    // the R stack (mpqr_create: allocations, default-stream memsets, stream creation) pass 1 stopped for good, in about one run of
    // three.  [B200, r2t] tests/test_gpu_parity_large.py::test_c5_full_size_vs_lapack: two panel_chain_kernel clusters resident,
    // nothing else dispatched, host in build_plan's cudaDeviceSynchronize (cuda-gdb listing: profiles/r2_tsqr_hang_cuda_gdb.txt).
    // The trigger was not isolated (a single handle takes cuMemAlloc in flight without harm, tools/alloc_hazard.py), so nothing
    // that waits runs here at all; stream-ordered lanes are not slower (23.7 ms at 8 lanes against 24.4 ms with the chain).
    const unsigned flags = MPQR_FP32 | (P->with_q ? MPQR_KEEP_WY : 0u) | MPQR_STREAM_ORDERED;
    const long ldp = P->ldp;
    auto fail_alloc = [&]() { set_error("tsqr: device allocation failed"); cudaGetLastError(); return MPQR_ENOMEM; };
    if (cudaEventCreateWithFlags(&P->ev_start, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&P->ev_tree, cudaEventDisableTiming) != cudaSuccess) { set_error("tsqr: event creation failed"); return MPQR_ECUDA; }
    if (nblk > 1 && cudaMalloc(&P->rstack, (size_t)nblk * n * ldp * sizeof(float)) != cudaSuccess) return fail_alloc();
    if (nblk > 1 && P->with_q && cudaMalloc(&P->qstack, (size_t)nblk * n * ldp * sizeof(float)) != cudaSuccess) return fail_alloc();
    P->ld32 = round_up(n, 4);   // = the FP32 driver's ld32 with MPQR_KEEP_WY (kmax = n for every block)
    if (nblk > 1 && P->with_q) {
        if (cudaMalloc(&P->Yall, (size_t)nblk * hrows * P->ld32 * sizeof(float)) != cudaSuccess) return fail_alloc();
        if (cudaMalloc(&P->Wall, (size_t)nblk * hrows * P->ld32 * sizeof(float)) != cudaSuccess) return fail_alloc();
    }
    for (int l = 0; l < P->NL; ++l) {
        Lane& L = P->lanes[l];
        if (cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking) != cudaSuccess) { set_error("tsqr: stream creation failed"); return MPQR_ECUDA; }
        if (cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming) != cudaSuccess) { set_error("tsqr: event creation failed"); return MPQR_ECUDA; }
        if (cudaMalloc(&L.work, (size_t)(hrows + 1) * ldp * sizeof(float)) != cudaSuccess) return fail_alloc();
        MPQR_TRY(mpqr_create(&L.hb, (int)hrows, n, r, 0, flags));
        if (P->hlast != hrows && (int)((nblk - 1) % P->NL) == l) MPQR_TRY(mpqr_create(&L.hl, (int)P->hlast, n, r, 0, flags));
    }
    // mpqr_create clears its workspaces on the legacy default stream; the lanes are non-blocking streams
    MPQR_CUDA(cudaDeviceSynchronize());
    return MPQR_OK;
}

int acquire_plan(long m, int n, bool with_q, Plan** out) {
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    int dev = 0;
    MPQR_CUDA(cudaGetDevice(&dev));
    const int NLwant = tsqr_lanes(1L << 30);
    std::lock_guard<std::mutex> lk(g_plans_mu);
    for (size_t i = 0; i < g_plans.size(); ++i) {
        Plan* P = g_plans[i];
        if (!P->busy && P->dev == dev && P->m == m && P->n == n && P->with_q == with_q && P->NL == (NLwant < P->nblk ? NLwant : (int)P->nblk)) {
            g_plans.erase(g_plans.begin() + i);  // most recently used at the back
            g_plans.push_back(P);
            P->busy = true;
            *out = P;
            return MPQR_OK;
        }
    }
    for (size_t i = 0; g_plans.size() >= kMaxPlans && i < g_plans.size(); ++i)
        if (!g_plans[i]->busy) { delete g_plans[i]; g_plans.erase(g_plans.begin() + i); break; }
    Plan* P = new Plan();
    P->dev = dev; P->m = m; P->n = n; P->with_q = with_q;
    int rc = build_plan(P, di);
    if (rc != MPQR_OK) { delete P; return rc; }
    P->busy = true;
    g_plans.push_back(P);
    *out = P;
    return MPQR_OK;
}

void release_plan(Plan* P) {
    std::lock_guard<std::mutex> lk(g_plans_mu);
    P->busy = false;
}

int run_plan(Plan* P, const float* dA, long lda, float* dQ, long ldq, float* dR, long ldr, cudaStream_t st) {
    const long nblk = P->nblk, hrows = P->hrows, hlast = P->hlast, ldp = P->ldp;
    const int n = P->n, NL = P->NL, budget = P->budget;
    auto& lanes = P->lanes;
    int rc = MPQR_OK;
    MPQR_CUDA(cudaEventRecord(P->ev_start, st));
    for (auto& L : lanes) MPQR_CUDA(cudaStreamWaitEvent(L.s, P->ev_start, 0));
    // Pass 1: every block is factored once; R_b goes to the stack, the block's reflectors into its slice of Yall / Wall
    // (the handle's own Y32 / W32 pointers are redirected for the call: same layout, ld32).
    for (long b = 0; b < nblk && rc == MPQR_OK; ++b) {
        Lane& L = lanes[b % NL];
        SmBudget sb(budget);
        const long rows = (b == nblk - 1) ? hlast : hrows;
        mpqr_handle* h = (rows == hrows) ? L.hb : L.hl;
        copy_block_kernel<<<grid_of(rows * n), 256, 0, L.s>>>(dA + (size_t)b * hrows * lda, lda, L.work, ldp, rows, n);
        float *y0 = h->Y32, *w0 = h->W32;
        if (P->Yall) {
            if (h->ld32 != P->ld32) { set_error("tsqr: internal layout mismatch"); return MPQR_ESTATE; }
            h->Y32 = P->Yall + (size_t)b * hrows * P->ld32;
            h->W32 = P->Wall + (size_t)b * hrows * P->ld32;
        }
        rc = mpqr_factor_device(h, L.work, ldp, L.s);
        const float *yb = h->Y32, *wb = h->W32;
        h->Y32 = y0; h->W32 = w0;
        if (rc) break;
        if (nblk == 1) {
            extract_r_kernel<<<grid_of((long)n * n), 256, 0, L.s>>>(L.work, ldp, dR, ldr, n);
            if (dQ) rc = form_thin_q(h, yb, wb, dQ, ldq, nullptr, 0, L.s);
        } else {
            extract_r_kernel<<<grid_of((long)n * n), 256, 0, L.s>>>(L.work, ldp, P->rstack + (size_t)b * n * ldp, ldp, n);
        }
    }
    // join the lanes on the caller's stream
    for (auto& L : lanes) { cudaEventRecord(L.done, L.s); cudaStreamWaitEvent(st, L.done, 0); }
    if (rc != MPQR_OK || nblk == 1) return rc;
    // factor the stacked R's ((nblk*n) x n) — recursion handles a tall stack
    MPQR_TRY(mpqr_tsqr_device(P->rstack, ldp, nblk * (long)n, n, dQ ? P->qstack : nullptr, ldp, dR, ldr, st));
    if (!dQ) return MPQR_OK;
    // Pass 2: thin Q rows of block b = Q_b [Qstack[b*n:(b+1)*n, :]; 0]
    MPQR_CUDA(cudaEventRecord(P->ev_tree, st));
    for (auto& L : lanes) MPQR_CUDA(cudaStreamWaitEvent(L.s, P->ev_tree, 0));
    for (long b = 0; b < nblk && rc == MPQR_OK; ++b) {
        Lane& L = lanes[b % NL];
        SmBudget sb(budget);
        const long rows = (b == nblk - 1) ? hlast : hrows;
        mpqr_handle* h = (rows == hrows) ? L.hb : L.hl;
        rc = form_thin_q(h, P->Yall + (size_t)b * hrows * P->ld32, P->Wall + (size_t)b * hrows * P->ld32, dQ + (size_t)b * hrows * ldq, ldq,
                         P->qstack + (size_t)b * n * ldp, ldp, L.s);
    }
    for (auto& L : lanes) { cudaEventRecord(L.done, L.s); cudaStreamWaitEvent(st, L.done, 0); }
    return rc;
}

}  // namespace

extern "C" int mpqr_tsqr_device(const float* dA, long lda, long m, int n, float* dQ, long ldq, float* dR, long ldr,
                                void* stream) {
    if (!dA || !dR || n < 1 || m < n || lda < n || ldr < n || (dQ && ldq < n) || m > 0x7fffffffL) {
        set_error("mpqr_tsqr_device: bad arguments (needs m >= n)");
        return MPQR_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Plan* P = nullptr;
    MPQR_TRY(acquire_plan(m, n, dQ != nullptr, &P));
    int rc = run_plan(P, dA, lda, dQ, ldq, dR, ldr, st);
    // the plan's buffers are reused by the next call: the call is synchronous with respect to the host
    cudaError_t e = cudaStreamSynchronize(st);
    for (auto& L : P->lanes) { cudaError_t e2 = cudaStreamSynchronize(L.s); if (e == cudaSuccess) e = e2; }
    release_plan(P);
    if (rc == MPQR_OK && e != cudaSuccess) { set_error("tsqr: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; }
    return rc;
}

extern "C" int mpqr_tsqr_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_plans_mu);
    for (size_t i = 0; i < g_plans.size();) {
        if (!g_plans[i]->busy) { delete g_plans[i]; g_plans.erase(g_plans.begin() + i); }
        else ++i;
    }
    return MPQR_OK;
}
