// api.cu — C-ABI (include/mpqr.h) and the blocked Householder QR drivers.
//
// Host C++ only issues kernels on one stream; the matrix stays device-resident for the whole
// factorisation (the reference moves the full (m+1) x n matrix over PCIe twice per panel:
// Cuda/qr.cu:1082, :1215, and synchronises after every launch).
//
// Driver structure (replaces dev_mixed_precision_block_qr / dev_block_qr_wy panel loops,
// Cuda/qr.cu:1074-1219 / :980-1040; panel bounds tau = min(lam + r, n) as :1076):
//
//   FP32 path  : for each panel  [panel kernel] -> S = W^T A22 -> A22 -= Y S        (SIMT)
//   FP16/BF16  : two-level blocking.  Outer block of nb columns, inner panels of r:
//       [panel kernel]  -> in-block update with the panel's own (Y_p, W_p)   (K = r)
//                       -> WY accumulation W_p <- W_p - W_prev (Y_prev^T W_p)  (GVL 5.1.2 WY form,
//                          the reference's W recurrence Cuda/qr.cu:361-399 done blockwise)
//       far update  A[c0:, c1:] -= Y_blk (W_blk^T A[c0:, c1:])                 (K = nb)
//     so the FP32 master of the far trailing matrix is read and written once per nb columns
//     instead of once per r columns (SURVEY 7 "hard parts": K >= ~512 needed to be
//     tensor-bound rather than HBM-bound).
//   The trailing update uses the identity  Q_panel^T A = A - Y (W^T A)  with W = Y T
//   ("mind the transpose").
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>

#include "internal.h"

namespace mpqr {

static thread_local char g_err[512] = "";
thread_local int g_sm_budget = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int get_device_info(DeviceInfo* out) {
    static thread_local DeviceInfo cached;
    static thread_local int cached_dev = -1;
    int dev;
    MPQR_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        MPQR_CUDA(cudaDeviceGetAttribute(&cached.num_sms, cudaDevAttrMultiProcessorCount, dev));
        MPQR_CUDA(cudaDeviceGetAttribute(&cached.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        MPQR_CUDA(cudaDeviceGetAttribute(&cached.coop, cudaDevAttrCooperativeLaunch, dev));
        int major = 0;
        MPQR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        if (major != 10) {
            set_error("libmpqr is built for sm_100a only; device has compute capability major %d", major);
            return MPQR_ECUDA;
        }
        cached_dev = dev;
    }
    *out = cached;
    return MPQR_OK;
}

thread_local double g_host_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
bool g_host_prof_on = getenv("MPQR_HOST_TRACE") != nullptr;
static inline double host_now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
HostProfScope::HostProfScope(int c) : cat(c), t0(g_host_prof_on ? host_now_ms() : 0.0) {}
HostProfScope::~HostProfScope() {
    if (g_host_prof_on) g_host_prof[cat] += host_now_ms() - t0;
}

int func_attr_once(const void* func, cudaFuncAttribute attr, int value) {
    struct Key { const void* f; int a, dev; };
    static std::mutex mu;
    static std::vector<Key> done;
    int dev = 0;
    MPQR_CUDA(cudaGetDevice(&dev));
    static thread_local std::vector<Key> seen;   // lock-free fast path of the issuing thread
    for (const Key& k : seen)
        if (k.f == func && k.a == (int)attr && k.dev == dev) return MPQR_OK;
    seen.push_back({func, (int)attr, dev});
    std::lock_guard<std::mutex> lk(mu);
    for (const Key& k : done)
        if (k.f == func && k.a == (int)attr && k.dev == dev) return MPQR_OK;
    MPQR_CUDA(cudaFuncSetAttribute(func, attr, value));
    done.push_back({func, (int)attr, dev});
    return MPQR_OK;
}

// ------------------------------------------------------------------------ utility kernels
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_uniform_kernel(float* A, long lda, long n_total, long row0, long rows, long col0,
                                    long cols, uint64_t seed_mixed) {
    long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long i = idx / cols, j = idx - i * cols;
        uint64_t h = mix64(seed_mixed + (uint64_t)((row0 + i) * n_total + (col0 + j)));
        A[i * lda + j] = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
}

__device__ __forceinline__ void cvt16(float v, __half* d) { *d = __float2half_rn(v); }
__device__ __forceinline__ void cvt16(float v, __nv_bfloat16* d) { *d = __float2bfloat16_rn(v); }

template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, long lds, T* __restrict__ dst, long ldd, long rows, long cols) {
    long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long i = idx / cols, j = idx - i * cols;
        cvt16(src[i * lds + j], dst + i * ldd + j);
    }
}

__global__ void identity_kernel(float* Q, long ldq, int m) {
    long total = (long)m * m;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long i = idx / m, j = idx - i * m;
        Q[i * ldq + j] = (i == j) ? 1.f : 0.f;
    }
}

__global__ void zero16_kernel(uint16_t* dst, long ldd, long rows, long cols) {
    long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long i = idx / cols, j = idx - i * cols;
        dst[i * ldd + j] = 0;
    }
}

int grid_for(long total) {
    long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

int preload_util_kernels() {
    const void* fns[] = {(const void*)fill_uniform_kernel, (const void*)convert_kernel<__half>, (const void*)convert_kernel<__nv_bfloat16>,
                         (const void*)identity_kernel, (const void*)zero16_kernel};
    for (const void* f : fns) {
        cudaFuncAttributes fa;
        MPQR_CUDA(cudaFuncGetAttributes(&fa, f));
    }
    return MPQR_OK;
}

int fill_uniform(float* A, long lda, long n_total, long row0, long rows, long col0, long cols, uint64_t seed,
                 cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return MPQR_OK;
    // host-side twin of mix64 for the seed (oracle/mpqr_oracle.c:orc_uniform01)
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    fill_uniform_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(A, lda, n_total, row0, rows, col0, cols, z);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

int convert_f32_to_16(const float* src, long lds, void* dst, long ldd, long rows, long cols, int bf16,
                      cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return MPQR_OK;
    if (bf16) convert_kernel<__nv_bfloat16><<<grid_for(rows * cols), 256, 0, stream>>>(src, lds, (__nv_bfloat16*)dst, ldd, rows, cols);
    else convert_kernel<__half><<<grid_for(rows * cols), 256, 0, stream>>>(src, lds, (__half*)dst, ldd, rows, cols);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

int set_identity(float* Q, long ldq, int m, cudaStream_t stream) {
    identity_kernel<<<grid_for((long)m * m), 256, 0, stream>>>(Q, ldq, m);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

int fill_zero_16(void* dst, long ldd, long rows, long cols, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return MPQR_OK;
    zero16_kernel<<<grid_for(rows * cols), 256, 0, stream>>>((uint16_t*)dst, ldd, rows, cols);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

}  // namespace mpqr

using namespace mpqr;

using namespace mpqr;

namespace {


// ------------------------------------------------------------------ look-ahead resources
// Two green contexts (CUDA 12.4+ driver API, resolved at run time so that libmpqr.so does not
// link libcuda): a small cluster-capable SM partition for the panel chain and the rest of the
// device for the far trailing update.  Streams created in a green context only use its SMs, so
// the panel chain of outer block J+1 really runs WHILE block J's far update does.
struct GreenApi {
    CUresult (*DeviceGet)(CUdevice*, int);
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    CUresult (*GreenCtxDestroy)(CUgreenCtx);
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
    bool ok;
};
const GreenApi* green_api() {
    static GreenApi api{};
    static bool tried = false;
    if (!tried) {
        tried = true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        api.ok = get("cuDeviceGet", (void**)&api.DeviceGet) && get("cuDeviceGetDevResource", (void**)&api.DeviceGetDevResource) &&
                 get("cuDevSmResourceSplitByCount", (void**)&api.DevSmResourceSplitByCount) &&
                 get("cuDevResourceGenerateDesc", (void**)&api.DevResourceGenerateDesc) &&
                 get("cuGreenCtxCreate", (void**)&api.GreenCtxCreate) && get("cuGreenCtxDestroy", (void**)&api.GreenCtxDestroy) &&
                 get("cuGreenCtxStreamCreate", (void**)&api.GreenCtxStreamCreate);
        if (!api.ok) cudaGetLastError();
    }
    return &api;
}

// Best effort: on any failure the handle simply keeps the single-stream driver.
// `sizes`: panel-partition SM counts to prepare (each with the rest of the device as its update partition).
void overlap_init(mpqr_handle* h, const int* sizes, int nsizes) {
    const GreenApi* g = green_api();
    if (!g->ok) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    CUdevice cudev;
    CUdevResource full;
    if (g->DeviceGet(&cudev, dev) != CUDA_SUCCESS) return;
    if (g->DeviceGetDevResource(cudev, &full, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
    auto& o = h->ov;
    o.nsm_full = (int)full.sm.smCount;
    for (int k = 0; k < nsizes; ++k) {
        CUdevResource grp[1], rem;
        unsigned int n = 1;
        if (g->DevSmResourceSplitByCount(grp, &n, &full, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE,
                                         (unsigned)sizes[k]) != CUDA_SUCCESS || n < 1 || rem.sm.smCount < 16)
            continue;
        CUdevResourceDesc dP, dU;
        if (g->DevResourceGenerateDesc(&dP, &grp[0], 1) != CUDA_SUCCESS) continue;
        if (g->DevResourceGenerateDesc(&dU, &rem, 1) != CUDA_SUCCESS) continue;
        CUgreenCtx gP = nullptr, gU = nullptr;
        if (g->GreenCtxCreate(&gP, dP, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) continue;
        if (g->GreenCtxCreate(&gU, dU, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { g->GreenCtxDestroy(gP); continue; }
        CUstream sP = nullptr, sP2 = nullptr, sP3 = nullptr, sU = nullptr;
        if (g->GreenCtxStreamCreate(&sP, gP, CU_STREAM_NON_BLOCKING, -1) != CUDA_SUCCESS ||
            g->GreenCtxStreamCreate(&sP2, gP, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
            g->GreenCtxStreamCreate(&sP3, gP, CU_STREAM_NON_BLOCKING, -1) != CUDA_SUCCESS ||
            g->GreenCtxStreamCreate(&sU, gU, CU_STREAM_NON_BLOCKING, -1) != CUDA_SUCCESS) {
            if (sP) cudaStreamDestroy((cudaStream_t)sP);
            if (sP2) cudaStreamDestroy((cudaStream_t)sP2);
            if (sP3) cudaStreamDestroy((cudaStream_t)sP3);
            g->GreenCtxDestroy(gP); g->GreenCtxDestroy(gU);
            continue;
        }
        mpqr_handle::Overlap::Pair pr;
        pr.gP = gP; pr.gU = gU; pr.sP = (cudaStream_t)sP; pr.sP2 = (cudaStream_t)sP2; pr.sP3 = (cudaStream_t)sP3; pr.sU = (cudaStream_t)sU;
        pr.nsmP = (int)grp[0].sm.smCount; pr.nsmU = (int)rem.sm.smCount;
        o.pairs.push_back(pr);
    }
    if (o.pairs.empty()) return;
    if (cudaStreamCreateWithFlags(&o.sF, cudaStreamNonBlocking) != cudaSuccess) { o.sF = nullptr; }
    if (cudaStreamCreateWithFlags(&o.sF2, cudaStreamNonBlocking) != cudaSuccess) { o.sF2 = nullptr; }
    if (cudaStreamCreateWithFlags(&o.sF3, cudaStreamNonBlocking) != cudaSuccess) { o.sF3 = nullptr; }
    o.ev_rest.resize(2 * (ceil_div(h->nb, h->r) + 1));
    for (auto& e : o.ev_rest) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    const int nblk = ceil_div(h->kmax, h->nb);
    o.ev_bp.resize(nblk); o.ev_fn.resize(nblk); o.ev_fr.resize(nblk);
    for (int i = 0; i < nblk; ++i) {
        cudaEventCreateWithFlags(&o.ev_bp[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&o.ev_fn[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&o.ev_fr[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&o.ev_start, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&o.ev_end, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&o.ev_accdone, cudaEventDisableTiming);
    o.ev_acc.resize(ceil_div(h->nb, h->r) + 1);
    for (auto& e : o.ev_acc) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    const char* trc = getenv("MPQR_TRACE");
    o.trace = trc && trc[0] == '1';
    if (o.trace) {
        o.tr.resize(nblk);
        for (auto& t : o.tr) { cudaEventCreate(&t.b0); cudaEventCreate(&t.b1); cudaEventCreate(&t.f0); cudaEventCreate(&t.f1); cudaEventCreate(&t.f2); t.psm = 0; }
    }
    o.on = o.sF != nullptr && o.sF2 != nullptr;
}

void overlap_destroy(mpqr_handle* h) {
    auto& o = h->ov;
    for (auto e : o.ev_bp) cudaEventDestroy(e);
    for (auto e : o.ev_fn) cudaEventDestroy(e);
    for (auto e : o.ev_fr) cudaEventDestroy(e);
    if (o.ev_start) cudaEventDestroy(o.ev_start);
    if (o.ev_end) cudaEventDestroy(o.ev_end);
    if (o.ev_accdone) cudaEventDestroy(o.ev_accdone);
    for (auto e : o.ev_acc) cudaEventDestroy(e);
    if (o.sF) cudaStreamDestroy(o.sF);
    if (o.sF2) cudaStreamDestroy(o.sF2);
    if (o.sF3) cudaStreamDestroy(o.sF3);
    for (auto e : o.ev_rest) cudaEventDestroy(e);
    const GreenApi* g = green_api();
    for (auto& pr : o.pairs) {
        if (pr.sP) cudaStreamDestroy(pr.sP);
        if (pr.sP2) cudaStreamDestroy(pr.sP2);
        if (pr.sP3) cudaStreamDestroy(pr.sP3);
        if (pr.sU) cudaStreamDestroy(pr.sU);
        if (g->ok) { g->GreenCtxDestroy((CUgreenCtx)pr.gP); g->GreenCtxDestroy((CUgreenCtx)pr.gU); }
    }
    o = mpqr_handle::Overlap();
}

// Cost model of the look-ahead driver (B200 measurements of round 2: profiles/r2_partition_sweep.txt).  Times in ms.
// Block phase of one outer block on a P-SM partition = latency of its panel chain + SM-time of everything else in it
// (side updates, finalize, Gram / T / W, in-block tensor-core updates) on the P - 16 SMs the cluster leaves free:
//   chain kernel of one r = 128 panel:  0.290 (D > 16384: 8 rows per thread)  /  0.140 + 0.0046 * D/1024  (D <= 16384)
//   rest: 3.81e-3 * D SM-ms per 8 panels   (fitted: D ~ 31744 -> 9.8 / 6.4 / 4.75 / 3.35 ms on 32 / 48 / 64 / 148 SMs)
// Far update: 4 D N' kb flops at ~8.4 TFLOP/s per SM (CTA-pair GEMMs: 4.26e12 flop in 6.0 ms on 84 SMs).
double model_bp_ms(const mpqr_handle* h, int c0, int c1, int sms) {
    double lat = 0, work = 0;
    for (int lam = c0; lam < c1; lam += h->r) {
        const double D = h->m - lam;
        lat += (D <= 16384) ? 0.140 + 0.0046 * D / 1024.0 : 0.290;
        work += 3.81e-3 * D / 8.0;
    }
    const double f = h->r / 128.0 < 0.25 ? 0.25 : h->r / 128.0;
    const int free_sms = sms > 24 ? sms - 16 : 8;
    return f * (lat + work / free_sms);
}
double model_far_ms(const mpqr_handle* h, int c0, int c1, int ncols, int sms) {
    const double flops = 4.0 * (double)(h->m - c0) * (double)ncols * (double)(c1 - c0);
    return flops / (8.4e12 * sms) * 1e3;
}

int factor_fp32(mpqr_handle* h, float* A, long lda, cudaStream_t st) {
    const int m = h->m, n = h->n, r = h->r;
    for (int lam = 0, p = 0; lam < h->kmax; lam += r, ++p) {
        const int pw = (lam + r < h->kmax) ? r : h->kmax - lam;
        const int tau = lam + pw, D = m - lam, nt = n - tau;
        float *Y, *W;
        long ld;
        if (h->keep_wy) {
            Y = h->Y32 + (size_t)lam * h->ld32 + lam;
            W = h->W32 + (size_t)lam * h->ld32 + lam;
            ld = h->ld32;
        } else {
            Y = h->Y32;
            W = h->W32;
            ld = r;
        }
        PanelArgs a{};
        a.A = A; a.lda = lda; a.m = m; a.n = n; a.lam = lam; a.acol = lam; a.pw = pw; a.blk_row0 = lam;
        a.Y32 = Y; a.W32 = W; a.ld32 = ld;
        a.T = h->T + (size_t)p * r * r; a.ldt = r;
        a.sync_ws = h->sync_ws; a.host_ctr = &h->sync_ctr; a.scratch = h->scratch; a.scratch_rows = h->scratch_rows;
        a.ws = h->panel_ws; a.ws_rows = h->panel_ws_rows; a.prof = h->prof ? &h->hook : nullptr;
        a.chain_side = h->no_chain ? nullptr : h->chain_side; a.chain_flags = h->chain_flags; a.chain_ctr = &h->chain_ctr;
        a.chain_last_far = &h->chain_last_far; a.chain_buf = p & 1;
        PROF(0, 4.0 * D * pw * pw, 8.0 * D * pw, launch_panel(a, st, &h->launches));
        if (nt > 0) {
            float* A22 = A + (size_t)lam * lda + tau;
            PROF(1, 2.0 * pw * nt * D, 4.0 * D * (pw + nt), sgemm_tn(W, ld, A22, lda, h->S32, h->lds32, pw, nt, D, st, &h->launches));
            PROF(2, 2.0 * D * nt * pw, 8.0 * D * nt, sgemm_nn_sub(Y, ld, h->S32, h->lds32, A22, lda, D, nt, pw, st, &h->launches));
        }
    }
    return MPQR_OK;
}

// per-chunk streams (in the update partition of pair `k`), events and GEMM scratch of the streamed-input schedule
int arrival_streams(mpqr_handle* h, int k) {
    auto& ar = h->arr;
    const GreenApi* g = green_api();
    const size_t nch = ar.c0.size();
    if (ar.cs.size() < nch) { ar.cs.resize(nch, nullptr); ar.cev.resize(nch, nullptr); ar.cS32.resize(nch, nullptr); ar.cS16.resize(nch, nullptr); }
    ar.cev_set.assign(nch, 0);
    for (size_t q = 1; q < nch; ++q) {
        if (!ar.cs[q]) {
            CUstream sq = nullptr;
            if (!g->ok || g->GreenCtxStreamCreate(&sq, (CUgreenCtx)h->ov.pairs[k].gU, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
                set_error("streamed input: stream creation failed");
                return MPQR_ECUDA;
            }
            ar.cs[q] = (cudaStream_t)sq;
            MPQR_CUDA(cudaEventCreateWithFlags(&ar.cev[q], cudaEventDisableTiming));
        }
        const size_t wdt = (size_t)round_up(ar.c1[q] - ar.c0[q], 8);
        if (!ar.cS32[q]) {
            MPQR_TRY(dev_alloc(h, (void**)&ar.cS32[q], (size_t)h->sk * h->lds32 * sizeof(float)));
            MPQR_TRY(dev_alloc(h, &ar.cS16[q], (size_t)h->sk * h->lds16 * 2));
        }
        (void)wdt;
    }
    return MPQR_OK;
}

int factor_16(mpqr_handle* h, float* A, long lda, cudaStream_t st) {
    const int m = h->m, n = h->n, nb = h->nb;
    const int bf = h->prec == 2;
    if ((lda & 3) || ((uintptr_t)A & 15)) {
        set_error("factor: the tensor-core path needs lda %% 4 == 0 and a 16-byte aligned dA (lda=%ld)", lda);
        return MPQR_EINVAL;
    }
    const int nblk = ceil_div(h->kmax, nb);
    // streamed input (host drop-in): needs the look-ahead driver and every block's W (catch-up of late chunks)
    const bool arriving = h->arr.on && h->ov.on && !h->ov.pairs.empty() && nblk >= 3 && !h->prof && h->keep_wy;
    if (!arriving) {
        if (h->arr.on) {  // the plan fell back to the plain schedule: everything must be there first
            for (cudaEvent_t e : h->arr.ev) MPQR_CUDA(cudaStreamWaitEvent(st, e, 0));
        } else {
            // operand shadow of the whole matrix
            PROF(3, 0, 6.0 * m * n, convert_f32_to_16(A, lda, h->Ah, h->ldh, m, n, bf, st));
            h->launches += 1;
        }
    }
    auto ctx_of = [&](int b, int c0) {
        BlockCtx c{};
        c.A = A; c.lda = lda; c.acol0 = c0; c.Ah = h->Ah; c.ldh = h->ldh;
        c.Y16 = at16(h->Ah, h->ldh, c0, c0); c.ldy = h->ldh;  // Y lives in the shadow's dead columns
        void* wbuf = (h->W16b && (b & 1)) ? h->W16b : h->W16;
        c.W16 = h->keep_wy ? (void*)at16(h->W16, h->ldw16, c0, c0) : wbuf;
        c.ldw = h->ldw16;
        return c;
    };
    // columns [c0, c1) are final once block_phase(b) is done on `producer`: ship them to the host sink
    auto emit = [&](int b, int c0, int c1, cudaStream_t producer) -> int {
        if (!h->sink_host) return MPQR_OK;
        HostProfScope hp(6);
        if ((int)h->sink_ev.size() <= b) {
            const size_t old = h->sink_ev.size();
            h->sink_ev.resize(b + 1);
            for (size_t i = old; i < h->sink_ev.size(); ++i) cudaEventCreateWithFlags(&h->sink_ev[i], cudaEventDisableTiming);
        }
        MPQR_CUDA(cudaEventRecord(h->sink_ev[b], producer));
        MPQR_CUDA(cudaStreamWaitEvent(h->sink_stream, h->sink_ev[b], 0));
        MPQR_CUDA(cudaMemcpy2DAsync(h->sink_host + c0, h->sink_pitch, A + c0, (size_t)lda * sizeof(float),
                                    (size_t)(c1 - c0) * sizeof(float), (size_t)m + 1, cudaMemcpyDeviceToHost, h->sink_stream));
        return MPQR_OK;
    };
    if (!h->ov.on || nblk < 3 || h->prof) {
        for (int c0 = 0, b = 0; c0 < h->kmax; c0 += nb, ++b) {
            const int c1 = (c0 + nb < h->kmax) ? c0 + nb : h->kmax;
            BlockCtx c = ctx_of(b, c0);
            MPQR_TRY(block_phase(h, c, c0, c1, c1 == n, st));
            MPQR_TRY(emit(b, c0, c1, st));
            MPQR_TRY(far_update(h, c, c0, c1, c1, n - c1, st));
        }
        if (h->kmax < n) MPQR_TRY(emit(nblk, h->kmax, n, st));  // wide matrices: the columns right of the last reflector
        return MPQR_OK;
    }
    // Look-ahead schedule.  Interval b = { far_next(b), far_rest(b) on an update partition ; block_phase(b+1) on
    // the matching panel partition }: the next outer block's columns are updated first (far_next) so that its
    // panel chain runs WHILE the rest of the trailing matrix is still being updated.  Per interval the
    // partition pair (or the whole device, serially) is chosen by the cost model above; everything is
    // ordered by events, so consecutive intervals may use different partitions.
    //   bp(b+1) waits fn(b);  fn(b) waits bp(b) and fr(b-1);  fr(b) follows fn(b) in its stream.
    auto& o = h->ov;
    MPQR_CUDA(cudaEventRecord(o.ev_start, st));
    const char* fix = getenv("MPQR_PANEL_SMS");
    const int fixed_sms = fix ? atoi(fix) : 0;
    // interval b-1 decided where block_phase(b) runs; block 0 runs on the whole device
    cudaStream_t s_bp = o.sF, s_bp2 = o.sF2, s_bp3 = o.sF3, s_uprev = nullptr;
    int nsm_bp = o.nsm_full, nsm_uprev = 0;
    const bool inblock_la = (h->r % 8) == 0;
    MPQR_CUDA(cudaStreamWaitEvent(s_bp, o.ev_start, 0));
    // Arrival-aware far updates (streamed host input).  The columns form a NEAR range [.., jnear), updated by the regular
    // far update of every interval on the update stream, and DISTANT chunks, each with its own stream in the update
    // partition: a distant chunk waits for its arrival once and then takes block b at every interval b until it joins the
    // near range, two blocks before the chain reaches it (its event orders the hand-over).  The backlog a late chunk has
    // to catch up with therefore runs on its own stream when the data lands -- never in front of the columns the panel
    // chain needs next -- and the host needs no clock.  One (panel, update) partition pair is used throughout.
    auto& ar = h->arr;
    size_t nnear = arriving ? 1 : 0;
    int jnear = arriving ? ar.c1[0] : n;
    int arr_pair = -1;
    if (arriving) {
        constexpr int arr_sms = 64;   // panel partition of the streamed schedule ([B200, r2v] 80 SMs: 191 ms against 167 ms)
        for (size_t k = 0; k < o.pairs.size(); ++k)
            if (arr_pair < 0 || abs(o.pairs[k].nsmP - arr_sms) < abs(o.pairs[arr_pair].nsmP - arr_sms)) arr_pair = (int)k;
        MPQR_TRY(arrival_streams(h, arr_pair));
        MPQR_CUDA(cudaStreamWaitEvent(s_bp, ar.ev[0], 0));
        for (size_t q = 1; q < ar.c0.size(); ++q) MPQR_CUDA(cudaStreamWaitEvent(ar.cs[q], ar.ev[q], 0));
    }
    // chunks whose first column lies left of `upto` join the near range (before the update stream touches them)
    auto promote = [&](int upto, bool all, cudaStream_t s_u) -> int {
        while (nnear < ar.c0.size() && (all || ar.c0[nnear] < upto)) {
            MPQR_CUDA(cudaStreamWaitEvent(s_u, ar.ev[nnear], 0));
            if (ar.cev_set[nnear]) MPQR_CUDA(cudaStreamWaitEvent(s_u, ar.cev[nnear], 0));  // blocks 0 .. b-1 applied on its own stream
            jnear = ar.c1[nnear];
            ++nnear;
        }
        return MPQR_OK;
    };
    for (int c0 = 0, b = 0; c0 < h->kmax; c0 += nb, ++b) {
        const int c1 = (c0 + nb < h->kmax) ? c0 + nb : h->kmax;
        BlockCtx c = ctx_of(b, c0);
        // In-block look-ahead: the part of every in-block update that is not the next panel's columns runs on a
        // second stream of the same partition, next to the next panel's register-block kernels.
        c.chain_side = s_bp3;
        if (inblock_la) {
            c.rest_stream = s_bp2; c.rest_S32 = h->S32r; c.rest_S16 = h->S16r; c.rest_ev = o.ev_rest.data();
            // leave the cluster (16 SMs) and a share for its side updates free
            constexpr int rest_keep = 32;   // SMs kept free of rest-stream GEMMs ([B200, r2p] 0 / 32 / 48: 114.1 / 114.2 / 114.4 ms)
            c.rest_sms = nsm_bp - rest_keep >= 24 ? nsm_bp - rest_keep : (nsm_bp >= 40 ? nsm_bp - 16 : 0);
            // (the previous block's last rest event completed before fn(b-1), which this block waits for)
        }
        // WY accumulation of this block (only the far update needs it): on the update partition of the previous
        // interval (it idles once its far update is done); else behind the in-block rest updates on their stream
        const bool defer_u = s_uprev && s_uprev != s_bp;
        cudaStream_t s_acc = defer_u ? s_uprev : (inblock_la ? s_bp2 : nullptr);
        if (s_acc) {
            c.acc_stream = s_acc;
            c.acc_sms = defer_u ? (nsm_uprev == o.nsm_full ? 0 : nsm_uprev) : (c.rest_sms > 0 ? c.rest_sms : (nsm_bp == o.nsm_full ? 0 : nsm_bp));
            c.acc_S32 = defer_u ? h->S32u : h->S32r; c.acc_S16 = defer_u ? h->S16u : h->S16r; c.acc_ev = o.ev_acc.data();
        }
        const bool defer_acc = s_acc != nullptr;
        if (b > 0) MPQR_CUDA(cudaStreamWaitEvent(s_bp, o.ev_fn[b - 1], 0));
        if (o.trace) { cudaEventRecord(o.tr[b].b0, s_bp); o.tr[b].psm = nsm_bp; }
        {
            SmBudget budget(nsm_bp == o.nsm_full ? 0 : nsm_bp);
            MPQR_TRY(block_phase(h, c, c0, c1, c1 == n, s_bp));
        }
        if (o.trace) cudaEventRecord(o.tr[b].b1, s_bp);
        MPQR_CUDA(cudaEventRecord(o.ev_bp[b], s_bp));
        if (defer_acc) MPQR_CUDA(cudaEventRecord(o.ev_accdone, s_acc));
        MPQR_TRY(emit(b, c0, c1, s_bp));
        // arrival mode: near range after this interval's promotions (the next two blocks must be in it; everything at the
        // last reflector block) -- only its width is needed here, for the cost model
        const bool last_blk = c1 >= h->kmax;
        int jplan = jnear;
        if (arriving)
            for (size_t q = nnear; q < ar.c0.size() && (last_blk || ar.c0[q] < c1 + 2 * nb); ++q) jplan = ar.c1[q];
        const int nfar = jplan - c1;
        // ---- choose the partition of interval b
        const int nnext = nfar < nb ? nfar : nb;
        const int c2 = (c1 + nb < h->kmax) ? c1 + nb : h->kmax;
        const bool has_next = c1 < h->kmax;
        int best = -1;  // -1: whole device, serial
        // (arrival mode: the distant range is updated next to the near one, and a partition is always used so that late
        //  far work never runs in front of the panel chain)
        const int nmodel = nfar - nnext;
        if (has_next && (nfar - nnext > 0 || arriving)) {
            double best_t = arriving ? 1e30 : model_bp_ms(h, c1, c2, o.nsm_full) + model_far_ms(h, c0, c1, nfar - nnext, o.nsm_full);
            for (size_t k = 0; k < o.pairs.size(); ++k) {
                if (arriving && (int)k != arr_pair) continue;   // the distant chunks' streams live in this pair's update partition
                if (fixed_sms > 0 && !arriving && o.pairs[k].nsmP != fixed_sms) continue;
                const double tb = 1.05 * model_bp_ms(h, c1, c2, o.pairs[k].nsmP);
                const double tf = model_far_ms(h, c0, c1, nmodel, o.pairs[k].nsmU);
                const double t = tb > tf ? tb : tf;
                if (t < best_t || (fixed_sms > 0 && best < 0)) { best_t = t; best = (int)k; }
            }
        }
        cudaStream_t s_u = best >= 0 ? o.pairs[best].sU : o.sF;
        const int nsm_u = best >= 0 ? o.pairs[best].nsmU : o.nsm_full;
        if (nfar > 0) {
            BlockCtx cu = c;
            cu.S32 = h->S32u; cu.S16 = h->S16u;
            cu.acc_stream = nullptr;
            SmBudget budget(nsm_u == o.nsm_full ? 0 : nsm_u);
            MPQR_CUDA(cudaStreamWaitEvent(s_u, o.ev_bp[b], 0));
            if (defer_acc) MPQR_CUDA(cudaStreamWaitEvent(s_u, o.ev_accdone, 0));
            if (b > 0) MPQR_CUDA(cudaStreamWaitEvent(s_u, o.ev_fr[b - 1], 0));
            if (o.trace) cudaEventRecord(o.tr[b].f0, s_u);
            if (arriving) {
                // near range: what the next block phase needs first, then the chunks two blocks ahead, then the rest
                MPQR_TRY(promote(c1 + nnext, false, s_u));
                MPQR_TRY(far_update(h, cu, c0, c1, c1, nnext, s_u));
                if (o.trace) cudaEventRecord(o.tr[b].f1, s_u);
                MPQR_CUDA(cudaEventRecord(o.ev_fn[b], s_u));
                MPQR_TRY(promote(c1 + 2 * nb, last_blk, s_u));
                MPQR_TRY(far_update(h, cu, c0, c1, c1 + nnext, jnear - c1 - nnext, s_u));
                if (o.trace) cudaEventRecord(o.tr[b].f2, s_u);
                MPQR_CUDA(cudaEventRecord(o.ev_fr[b], s_u));
                // distant chunks: block b each, on their own streams
                for (size_t q = nnear; q < ar.c0.size(); ++q) {
                    BlockCtx cd = c;
                    cd.S32 = ar.cS32[q]; cd.S16 = ar.cS16[q];
                    cd.acc_stream = nullptr;
                    SmBudget budget(o.pairs[arr_pair].nsmU / 2);   // far_next must always find free SMs in the update partition
                    MPQR_CUDA(cudaStreamWaitEvent(ar.cs[q], o.ev_bp[b], 0));
                    if (defer_acc) MPQR_CUDA(cudaStreamWaitEvent(ar.cs[q], o.ev_accdone, 0));
                    MPQR_TRY(far_update(h, cd, c0, c1, ar.c0[q], ar.c1[q] - ar.c0[q], ar.cs[q]));
                    MPQR_CUDA(cudaEventRecord(ar.cev[q], ar.cs[q]));
                    ar.cev_set[q] = 1;
                }
            } else {
            MPQR_TRY(far_update(h, cu, c0, c1, c1, nnext, s_u));
            if (o.trace) cudaEventRecord(o.tr[b].f1, s_u);
            MPQR_CUDA(cudaEventRecord(o.ev_fn[b], s_u));
            MPQR_TRY(far_update(h, cu, c0, c1, c1 + nnext, nfar - nnext, s_u));
            if (o.trace) cudaEventRecord(o.tr[b].f2, s_u);
            MPQR_CUDA(cudaEventRecord(o.ev_fr[b], s_u));
            }
        } else {
            if (defer_acc) MPQR_CUDA(cudaStreamWaitEvent(s_bp, o.ev_accdone, 0));
            MPQR_CUDA(cudaEventRecord(o.ev_fn[b], s_bp));
            MPQR_CUDA(cudaEventRecord(o.ev_fr[b], s_bp));
        }
        // where the next block_phase runs
        s_bp = best >= 0 ? o.pairs[best].sP : o.sF;
        s_bp2 = best >= 0 ? o.pairs[best].sP2 : o.sF2;
        s_bp3 = best >= 0 ? o.pairs[best].sP3 : o.sF3;
        nsm_bp = best >= 0 ? o.pairs[best].nsmP : o.nsm_full;
        s_uprev = best >= 0 ? s_u : nullptr;
        nsm_uprev = nsm_u;
    }
    // join: the last ev_fr / ev_bp cover everything (fr(b) follows fn(b); bp(last) waited fn(last-1))
    MPQR_CUDA(cudaStreamWaitEvent(st, o.ev_fr[nblk - 1], 0));
    MPQR_CUDA(cudaStreamWaitEvent(st, o.ev_bp[nblk - 1], 0));
    for (int b = 0; b + 1 < nblk; ++b) MPQR_CUDA(cudaStreamWaitEvent(st, o.ev_fr[b], 0));
    if (arriving)
        for (size_t q = 1; q < ar.c0.size(); ++q)
            if (ar.cev_set[q]) MPQR_CUDA(cudaStreamWaitEvent(st, ar.cev[q], 0));
    if (h->kmax < n) MPQR_TRY(emit(nblk, h->kmax, n, st));
    return MPQR_OK;
}

}  // namespace

namespace mpqr {

int dev_alloc(mpqr_handle* h, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return MPQR_ENOMEM;
    }
    h->allocs.push_back(*p);
    return MPQR_OK;
}

int block_phase(mpqr_handle* h, const BlockCtx& c, int c0, int c1, int end_is_matrix_end, cudaStream_t st) {
    const int m = h->m, n = h->n, r = h->r;
    const int bf = h->prec == 2;
    const int Dblk = m - c0;
    float* S32 = c.S32 ? c.S32 : h->S32;
    void* S16 = c.S16 ? c.S16 : h->S16;
    int last_rest = -1;
    std::vector<char> rest_rec((size_t)ceil_div(c1 - c0, r) + 1, 0);
    for (int lam = c0; lam < c1; lam += r) {
        const int p = lam / r;
        const int pw = (lam + r < c1) ? r : c1 - lam;
        const int tau = lam + pw, D = m - lam, jc = lam - c0;
        PanelArgs a{};
        a.A = c.A; a.lda = c.lda; a.m = m; a.n = n; a.lam = lam; a.acol = c.acol0 + jc; a.pw = pw; a.blk_row0 = c0;
        a.W32 = h->Wblk32 + jc; a.ld32 = h->ldwb;  // FP32 master of W (Y32 not needed)
        a.Y16 = (char*)c.Y16 + (size_t)jc * 2; a.ldy16 = c.ldy;
        a.W16 = (char*)c.W16 + (size_t)jc * 2; a.ldw16 = c.ldw;
        a.bf16 = bf;
        a.T = h->T + (size_t)p * r * r; a.ldt = r;
        a.sync_ws = h->sync_ws; a.host_ctr = &h->sync_ctr; a.scratch = h->scratch; a.scratch_rows = h->scratch_rows;
        a.ws = h->panel_ws; a.ws_rows = h->panel_ws_rows; a.prof = h->prof ? &h->hook : nullptr;
        a.chain_side = h->no_chain ? nullptr : (c.chain_side ? c.chain_side : h->chain_side); a.chain_flags = h->chain_flags; a.chain_ctr = &h->chain_ctr;
        a.chain_last_far = &h->chain_last_far;
        a.chain_buf = (int)(h->chain_panels++ & 1u);
        const int nin = c1 - tau;  // in-block trailing columns
        const int pidx = jc / r;
        const bool chain = panel_chain_ok(a);
        // the next panel's columns on the panel stream, the rest of the block on `rest_stream` (same SM partition) next
        // to the next panel's chain kernel
        const bool split = c.rest_stream && nin > r && (r % 8) == 0;
        // Merged Gram / next-panel product (PanelArgs::gs_ncols): the panel's T kernel leaves S = T^T Y^T A_next in S16, the
        // in-block update of the next panel's columns is one NN GEMM, and W = Y T (only the rest of the block, the WY
        // accumulation and the far update need it) moves to the rest stream.  Needs Y in the shadow's dead columns.
        const bool y_in_shadow = c.Y16 == (void*)at16(c.Ah, c.ldh, c0, c.acol0) && c.ldy == c.ldh;
        const bool merged = !h->prof && y_in_shadow && nin > 0 && (split || nin <= r) && (r % 8) == 0 && (pw % 8) == 0 &&
                            (h->lds16 % 8) == 0;
        if (merged) { a.gs_ncols = nin < r ? nin : r; a.gs_S16 = S16; a.gs_lds16 = h->lds16; a.defer_w = 1; }
        // earlier panels' updates of THIS panel's columns: panels <= p-2 through their rest updates (rest stream, in order);
        // panel p-1 through its in-block update on this stream, or through its side updates (flag inside a chain kernel,
        // the side event for any other kernel)
        if (pidx >= 2 && rest_rec[pidx - 2]) MPQR_CUDA(cudaStreamWaitEvent(st, c.rest_ev[2 * (pidx - 2) + 1], 0));
        PROF(0, 4.0 * D * pw * pw, 8.0 * D * pw, launch_panel(a, st, &h->launches));
        const int acol_tau = c.acol0 + jc + pw;
        // in-block update of the columns [tau + ofs, tau + ofs + nc):  S = W_p^T A ; A -= Y_p S (+ shadow)
        const void* Wpp = (char*)c.W16 + ((size_t)jc * c.ldw + jc) * 2;   // the panel's own W / Y, rows lam..
        const void* Ypp = (char*)c.Y16 + ((size_t)jc * c.ldy + jc) * 2;
        auto inblock = [&](int ofs, int nc, float* xS32, void* xS16, int pad_ok, cudaStream_t st) -> int {
            HostProfScope hp(2);
            int w16 = 0;
            PROF(1, 2.0 * pw * nc * D, tn_bytes(pw, nc, D),
                 tc_gemm_tn16(Wpp, c.ldw, at16(c.Ah, c.ldh, lam, acol_tau + ofs), c.ldh, xS32, h->lds32, xS16, h->lds16, &w16, pw, nc, D, bf, 1, st, &h->launches));
            if (!w16) {
                PROF(3, 0, 6.0 * pw * nc, convert_f32_to_16(xS32, h->lds32, xS16, h->lds16, pw, nc, bf, st));
                h->launches += 1;
            }
            PROF(2, 2.0 * D * nc * pw, nn_bytes(D, nc, pw),
                 tc_gemm_nn(Ypp, c.ldy, xS16, h->lds16, c.A + (size_t)lam * c.lda + acol_tau + ofs, c.lda,
                            at16(c.Ah, c.ldh, lam, acol_tau + ofs), c.ldh, D, nc, pw, bf, pad_ok, st, &h->launches));
            return MPQR_OK;
        };
        // the merged flow's in-block update of the next panel's columns: S is already in S16
        auto inblock_nn_only = [&](int nc, int pad_ok, cudaStream_t st) -> int {
            HostProfScope hp(2);
            return tc_gemm_nn(Ypp, c.ldy, S16, h->lds16, c.A + (size_t)lam * c.lda + acol_tau, c.lda, at16(c.Ah, c.ldh, lam, acol_tau), c.ldh,
                              D, nc, pw, bf, pad_ok, st, &h->launches);
        };
        if (c.rest_stream && nin > 0) { last_rest = 2 * pidx + 1; rest_rec[pidx] = true; }
        if (nin > 0 && !split) {
            if (c.rest_stream && pidx > 0) MPQR_CUDA(cudaStreamWaitEvent(st, c.rest_ev[2 * (pidx - 1) + 1], 0));
            if (merged) {
                MPQR_TRY(inblock_nn_only(nin, end_is_matrix_end, st));
                MPQR_TRY(panel_form_w(a, st, &h->launches));
            } else {
                MPQR_TRY(inblock(0, nin, S32, S16, end_is_matrix_end, st));
            }
            if (c.rest_stream) MPQR_CUDA(cudaEventRecord(c.rest_ev[2 * pidx + 1], st));
        } else if (nin > 0) {
            // the rest of the block on the second stream (it only needs this panel's Y, W and the previous rest)
            MPQR_CUDA(cudaEventRecord(c.rest_ev[2 * pidx], st));
            MPQR_CUDA(cudaStreamWaitEvent(c.rest_stream, c.rest_ev[2 * pidx], 0));
            // the next panel's columns were last written by the previous panel's rest update
            if (pidx > 0) MPQR_CUDA(cudaStreamWaitEvent(st, c.rest_ev[2 * (pidx - 1) + 1], 0));
            if (merged) MPQR_TRY(inblock_nn_only(r, 0, st));
            else MPQR_TRY(inblock(0, r, S32, S16, 0, st));
            // (holding the rest-of-block GEMMs back until the NEXT panel's cluster is resident -- they keep it from being placed for
            //  47-83 us -- was measured and dropped: the cluster starts sooner, but the GEMMs then run next to its side updates
            //  and slow them down by as much; [B200, r2p] 114.1-114.8 ms either way)
            {
                SmBudget rb(c.rest_sms > 0 ? c.rest_sms : g_sm_budget);
                if (merged) MPQR_TRY(panel_form_w(a, c.rest_stream, &h->launches));
                MPQR_TRY(inblock(r, nin - r, c.rest_S32, c.rest_S16, end_is_matrix_end, c.rest_stream));
            }
            MPQR_CUDA(cudaEventRecord(c.rest_ev[2 * pidx + 1], c.rest_stream));
        }
        if (jc > 0) {
            // WY accumulation: X = Y_prev^T W_p ; W_p -= W_prev X   (rows c0..m)
            const void* Wp = (char*)c.W16 + (size_t)jc * 2;
            cudaStream_t st_panel = st;
            float* aS32 = S32;
            void* aS16 = S16;
            if (c.acc_stream) {
                // W_p is modified in place: wait until this panel's in-block update has consumed it
                cudaEvent_t ev = c.acc_ev[jc / r];
                MPQR_CUDA(cudaEventRecord(ev, st_panel));
                MPQR_CUDA(cudaStreamWaitEvent(c.acc_stream, ev, 0));
                if (c.rest_stream && nin > 0) MPQR_CUDA(cudaStreamWaitEvent(c.acc_stream, c.rest_ev[2 * pidx + 1], 0));
                aS32 = c.acc_S32; aS16 = c.acc_S16;
            }
            {
                HostProfScope hp(3);
                cudaStream_t st = c.acc_stream ? c.acc_stream : st_panel;  // (PROF records on `st`)
                SmBudget budget(c.acc_stream ? c.acc_sms : g_sm_budget);
                int w16 = 0;
                PROF(1, 2.0 * jc * pw * Dblk, tn_bytes(jc, pw, Dblk),
                     tc_gemm_tn16(c.Y16, c.ldy, Wp, c.ldw, aS32, h->lds32, aS16, h->lds16, &w16, jc, pw, Dblk, bf, 1, st, &h->launches));
                if (!w16) {
                    PROF(3, 0, 6.0 * jc * pw, convert_f32_to_16(aS32, h->lds32, aS16, h->lds16, jc, pw, bf, st));
                    h->launches += 1;
                }
                PROF(2, 2.0 * Dblk * pw * jc, nn_bytes(Dblk, pw, jc),
                     tc_gemm_nn(c.W16, c.ldw, aS16, h->lds16, h->Wblk32 + jc, h->ldwb, (void*)Wp, c.ldw, Dblk, pw, jc, bf,
                                // a deferred accumulation may run after the NEXT panel produced its W: the zero spill of a
                                // clipped TMA store (16-byte granules) must not reach those columns
                                (c.acc_stream && (pw & 7)) ? 0 : 1, st, &h->launches));
            }
        }
    }
    // everything of this block is complete when the panel stream is (events of one stream complete in order)
    if (c.rest_stream && last_rest >= 0) MPQR_CUDA(cudaStreamWaitEvent(st, c.rest_ev[last_rest], 0));
    return MPQR_OK;
}

int far_update(mpqr_handle* h, const BlockCtx& c, int c0, int c1, int afar, int nfar, cudaStream_t st) {
    if (nfar <= 0) return MPQR_OK;
    HostProfScope hp(4);
    const int bf = h->prec == 2;
    const int Dblk = h->m - c0, kb = c1 - c0;
    float* S32 = c.S32 ? c.S32 : h->S32;
    void* S16 = c.S16 ? c.S16 : h->S16;
    int w16 = 0;
    PROF(1, 2.0 * kb * nfar * Dblk, tn_bytes(kb, nfar, Dblk),
         tc_gemm_tn16(c.W16, c.ldw, at16(c.Ah, c.ldh, c0, afar), c.ldh, S32, h->lds32, S16, h->lds16, &w16, kb, nfar, Dblk, bf, 1, st, &h->launches));
    if (!w16) {
        PROF(3, 0, 6.0 * kb * nfar, convert_f32_to_16(S32, h->lds32, S16, h->lds16, kb, nfar, bf, st));
        h->launches += 1;
    }
    PROF(2, 2.0 * Dblk * nfar * kb, nn_bytes(Dblk, nfar, kb),
         tc_gemm_nn(c.Y16, c.ldy, S16, h->lds16, c.A + (size_t)c0 * c.lda + afar, c.lda, at16(c.Ah, c.ldh, c0, afar), c.ldh,
                    Dblk, nfar, kb, bf, 1, st, &h->launches));
    return MPQR_OK;
}

}  // namespace mpqr

extern "C" {

const char* mpqr_last_error(void) { return g_err; }
const char* mpqr_version(void) { return "mpqr-b200 0.1 (sm_100a)"; }

int mpqr_create(mpqr_handle** out, int m, int n, int r, int nb, unsigned flags) {
    if (!out || m < 1 || n < 1 || r < 1) {
        set_error("mpqr_create: bad arguments m=%d n=%d r=%d", m, n, r);
        return MPQR_EINVAL;
    }
    unsigned prec = flags & MPQR_PRECISION_MASK;
    if (prec == 3) {
        set_error("mpqr_create: choose one of MPQR_FP16 / MPQR_BF16");
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    MPQR_TRY(chain_preload_all());   // (once per device)
    mpqr_handle* h = new mpqr_handle();
    h->m = m; h->n = n; h->flags = flags; h->prec = (int)prec;
    h->keep_wy = (flags & MPQR_KEEP_WY) != 0;
    h->no_chain = (flags & MPQR_STREAM_ORDERED) != 0;
    h->kmax = m < n ? m : n;
    h->r = r > kPanelMaxWidth ? kPanelMaxWidth : r;
    {
        // The packed result does not depend on the panel grouping beyond rounding, so a narrow caller r (the reference's
        // r = 16 / 32 / 64) is widened to a multiple of it near 128: fewer, better filled panels.
        // On by default (MPQR_MIN_R=0 keeps the caller's r): [B200] 2048^2 r = 32: 5.6 -> 3.7 ms, 4096 x 16384 r = 64: 8.2 -> 7.4 ms.
        const char* e = getenv("MPQR_MIN_R");
        const int min_r = e ? atoi(e) : kPanelMaxWidth;
        if (min_r > h->r && min_r <= kPanelMaxWidth) h->r = (min_r / h->r) * h->r;
    }
    if (h->r > h->kmax) h->r = h->kmax;
    if (prec == 0) {
        h->nb = h->r;
    } else {
        int want = nb > 0 ? nb : 1024;
        if (want < h->r) want = h->r;
        want = (want / h->r) * h->r;
        // keep at least ~2 outer blocks' worth of work; tiny problems use one level
        if (want > h->kmax) want = ceil_div(h->kmax, h->r) * h->r;
        h->nb = want;
    }
    h->npanels = ceil_div(h->kmax, h->r);
    int rc = MPQR_OK;
    do {
        if ((rc = dev_alloc(h, (void**)&h->sync_ws, panel_sync_ws_bytes()))) break;
        if (cudaMemset(h->sync_ws, 0, panel_sync_ws_bytes()) != cudaSuccess) { set_error("memset failed"); rc = MPQR_ECUDA; break; }
        // panel scratch only needed when a panel slice cannot live in shared memory
        if (m > 32768) {  // only the shared-memory fallback kernel (panel_legacy.cu) needs it
            h->scratch_rows = m;
            if ((rc = dev_alloc(h, (void**)&h->scratch, panel_scratch_bytes(m)))) break;
        }
        if ((rc = dev_alloc(h, (void**)&h->T, (size_t)h->npanels * h->r * h->r * sizeof(float)))) break;
        if ((rc = dev_alloc(h, (void**)&h->chain_flags, 64))) break;
        if (cudaMemset(h->chain_flags, 0, 64) != cudaSuccess) { set_error("memset failed"); rc = MPQR_ECUDA; break; }
        if (cudaStreamCreateWithFlags(&h->chain_side, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); rc = MPQR_ECUDA; break; }
        h->panel_ws_rows = m;
        if ((rc = dev_alloc(h, (void**)&h->panel_ws, panel_ws_bytes(m)))) break;
        const int wide = m > n ? m : n;
        h->sk = h->nb;
        h->lds32 = round_up(wide, 8);
        if ((rc = dev_alloc(h, (void**)&h->S32, (size_t)h->sk * h->lds32 * sizeof(float)))) break;
        if (prec == 0) {
            if (h->keep_wy) {
                h->ld32 = round_up(h->kmax, 4);
                if ((rc = dev_alloc(h, (void**)&h->Y32, (size_t)m * h->ld32 * sizeof(float)))) break;
                if ((rc = dev_alloc(h, (void**)&h->W32, (size_t)m * h->ld32 * sizeof(float)))) break;
            } else {
                h->ld32 = h->r;
                if ((rc = dev_alloc(h, (void**)&h->Y32, (size_t)m * h->r * sizeof(float)))) break;
                if ((rc = dev_alloc(h, (void**)&h->W32, (size_t)m * h->r * sizeof(float)))) break;
            }
        } else {
            h->ldh = round_up(n, 8);
            if ((rc = dev_alloc(h, &h->Ah, (size_t)m * h->ldh * 2))) break;
            h->ldw16 = h->keep_wy ? round_up(h->kmax, 8) : round_up(h->nb, 8);
            if ((rc = dev_alloc(h, &h->W16, (size_t)m * h->ldw16 * 2))) break;
            h->ldwb = round_up(h->nb, 4);
            if ((rc = dev_alloc(h, (void**)&h->Wblk32, (size_t)m * h->ldwb * sizeof(float)))) break;
            h->lds16 = h->lds32;
            if ((rc = dev_alloc(h, &h->S16, (size_t)h->sk * h->lds16 * 2))) break;
            // Look-ahead on two green-context SM partitions: on by default for problems with >= 4 outer
            // blocks (measured on B200, 32768^2: 181 ms serial, 159 ms with an 80-SM panel partition; a
            // small partition starves the panel chain's device-wide kernels: 280 ms at 16 SMs).
            const char* env = getenv("MPQR_OVERLAP");
            const int nblk_outer = ceil_div(h->kmax, h->nb);
            // (round 1, with the register-block look-ahead: 8192^2 19.3 -> 18.2 ms, 4096 x 16384 10.7 -> 9.4 ms, so from 4 blocks on)
            const bool want_overlap = env ? (env[0] == '1' && nblk_outer >= 3) : (nblk_outer >= 4);
            if (want_overlap) {
                const int sizes[5] = {32, 48, 64, 80, 112};
                overlap_init(h, sizes, 5);
                if (h->ov.on) {
                    if ((rc = dev_alloc(h, (void**)&h->S32u, (size_t)h->sk * h->lds32 * sizeof(float)))) break;
                    if ((rc = dev_alloc(h, (void**)&h->S32r, (size_t)h->sk * h->lds32 * sizeof(float)))) break;
                    if ((rc = dev_alloc(h, &h->S16r, (size_t)h->sk * h->lds16 * 2))) break;
                    if ((rc = dev_alloc(h, &h->S16u, (size_t)h->sk * h->lds16 * 2))) break;
                    if (!h->keep_wy && (rc = dev_alloc(h, &h->W16b, (size_t)m * h->ldw16 * 2))) break;
                }
            }
        }
    } while (0);
    if (rc != MPQR_OK) {
        mpqr_destroy(h);
        return rc;
    }
    *out = h;
    return MPQR_OK;
}

int mpqr_destroy(mpqr_handle* h) {
    if (!h) return MPQR_OK;
    mg_destroy(h->mg);
    for (auto sq : h->arr.cs) if (sq) cudaStreamDestroy(sq);   // (streams of a green context: before the contexts go)
    for (auto e : h->arr.cev) if (e) cudaEventDestroy(e);
    overlap_destroy(h);
    if (h->chain_side) cudaStreamDestroy(h->chain_side);
    for (auto e : h->arr.ev) cudaEventDestroy(e);
    if (h->arr.stream) cudaStreamDestroy(h->arr.stream);

    for (auto e : h->sink_ev) cudaEventDestroy(e);
    if (h->sink_stream) cudaStreamDestroy(h->sink_stream);
    for (void* p : h->allocs) cudaFree(p);
    for (auto& r : h->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : h->prof_pool) cudaEventDestroy(e);
    delete h;
    return MPQR_OK;
}

int mpqr_factor_device(mpqr_handle* h, float* dA, long lda, void* stream) {
    if (!h || !dA || lda < h->n) {
        set_error("mpqr_factor_device: bad arguments");
        return MPQR_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    h->launches = 0;
    int rc = h->prec == 0 ? factor_fp32(h, dA, lda, st) : factor_16(h, dA, lda, st);
    h->factored = (rc == MPQR_OK);
    return rc;
}

int mpqr_form_q_device(mpqr_handle* h, float* dQ, long ldq, void* stream) {
    if (!h || !dQ || ldq < h->m) {
        set_error("mpqr_form_q_device: bad arguments");
        return MPQR_EINVAL;
    }
    if (!h->factored || !h->keep_wy) {
        set_error("mpqr_form_q_device: needs MPQR_KEEP_WY and a completed mpqr_factor_device");
        return MPQR_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int m = h->m;
    h->launches = 0;
    MPQR_TRY(set_identity(dQ, ldq, m, st));
    h->launches += 1;
    if (h->prec == 0) {
        // Q <- Q_p Q = Q - W_p (Y_p^T Q), panels last to first; only Q[lam:, lam:] changes.
        for (int p = h->npanels - 1; p >= 0; --p) {
            const int lam = p * h->r;
            const int pw = (lam + h->r < h->kmax) ? h->r : h->kmax - lam;
            const int D = m - lam;
            float* Y = h->Y32 + (size_t)lam * h->ld32 + lam;
            float* W = h->W32 + (size_t)lam * h->ld32 + lam;
            float* Qs = dQ + (size_t)lam * ldq + lam;
            MPQR_TRY(sgemm_tn(Y, h->ld32, Qs, ldq, h->S32, h->lds32, pw, D, D, st, &h->launches));
            MPQR_TRY(sgemm_nn_sub(W, h->ld32, h->S32, h->lds32, Qs, ldq, D, D, pw, st, &h->launches));
        }
        return MPQR_OK;
    }
    const int bf = h->prec == 2;
    if ((ldq & 3) || ((uintptr_t)dQ & 15)) {
        set_error("form_q: the tensor-core path needs ldq %% 4 == 0 and a 16-byte aligned dQ");
        return MPQR_EINVAL;
    }
    if (!h->Qh) {
        h->ldqh = round_up(m, 8);
        MPQR_TRY(dev_alloc(h, &h->Qh, (size_t)m * h->ldqh * 2));
    }
    MPQR_TRY(convert_f32_to_16(dQ, ldq, h->Qh, h->ldqh, m, m, bf, st));
    h->launches += 1;
    const int nblk = ceil_div(h->kmax, h->nb);
    for (int b = nblk - 1; b >= 0; --b) {
        const int c0 = b * h->nb;
        const int c1 = (c0 + h->nb < h->kmax) ? c0 + h->nb : h->kmax;
        const int D = m - c0, kb = c1 - c0;
        const void* Yb = at16(h->Ah, h->ldh, c0, c0);
        const void* Wb = at16(h->W16, h->ldw16, c0, c0);
        MPQR_TRY(tc_gemm_tn(Yb, h->ldh, at16(h->Qh, h->ldqh, c0, c0), h->ldqh, h->S32, h->lds32, kb, D, D, bf, 1, st, &h->launches));
        MPQR_TRY(convert_f32_to_16(h->S32, h->lds32, h->S16, h->lds16, kb, D, bf, st));
        h->launches += 1;
        MPQR_TRY(tc_gemm_nn(Wb, h->ldw16, h->S16, h->lds16, dQ + (size_t)c0 * ldq + c0, ldq, at16(h->Qh, h->ldqh, c0, c0),
                            h->ldqh, D, D, kb, bf, 1, st, &h->launches));
    }
    return MPQR_OK;
}

int mpqr_get_panel_T(mpqr_handle* h, int panel, float* dT, int ldt, void* stream) {
    if (!h || !dT || panel < 0 || panel >= h->npanels || ldt < h->r) {
        set_error("mpqr_get_panel_T: bad arguments");
        return MPQR_EINVAL;
    }
    if (!h->factored) {
        set_error("mpqr_get_panel_T: factor first");
        return MPQR_ESTATE;
    }
    MPQR_CUDA(cudaMemcpy2DAsync(dT, (size_t)ldt * sizeof(float), h->T + (size_t)panel * h->r * h->r,
                                (size_t)h->r * sizeof(float), (size_t)h->r * sizeof(float), h->r,
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MPQR_OK;
}

int mpqr_num_panels(const mpqr_handle* h) { return h ? h->npanels : MPQR_EINVAL; }
int mpqr_effective_r(const mpqr_handle* h) { return h ? h->r : MPQR_EINVAL; }
int mpqr_effective_nb(const mpqr_handle* h) { return h ? h->nb : MPQR_EINVAL; }
long mpqr_last_launch_count(const mpqr_handle* h) { return h ? h->launches : MPQR_EINVAL; }

int mpqr_set_profiling(mpqr_handle* h, int on) {
    if (!h) return MPQR_EINVAL;
    for (auto& r : h->prof_recs) { h->prof_pool.push_back(r.e0); h->prof_pool.push_back(r.e1); }
    h->prof_recs.clear();
    for (int c = 0; c < MPQR_NUM_KERNEL_CLASSES; ++c) h->prof_flops[c] = h->prof_bytes[c] = 0;
    h->prof = on != 0;
    h->hook = {h, hook_begin, hook_end};
    return MPQR_OK;
}

int mpqr_get_profile(mpqr_handle* h, int cls, double* ms_total, long* launches, double* flops, double* bytes) {
    if (!h || cls < 0 || cls >= MPQR_NUM_KERNEL_CLASSES) return MPQR_EINVAL;
    double ms = 0;
    long cnt = 0;
    for (auto& r : h->prof_recs) {
        if (r.cls != cls) continue;
        MPQR_CUDA(cudaEventSynchronize(r.e1));
        float t = 0;
        MPQR_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms += t;
        ++cnt;
    }
    if (ms_total) *ms_total = ms;
    if (launches) *launches = cnt;
    if (flops) *flops = h->prof_flops[cls];
    if (bytes) *bytes = h->prof_bytes[cls];
    return MPQR_OK;
}

// The host drop-in keeps its plan (handle + device copies of A and Q) for the next call with the same shape on the
// same device: creating and above all FREEING ~12 GB of workspaces per call costs 30-330 ms at 32768^2 (measured on
// B200: cudaFree + green-context teardown), more than the factorisation.  mpqr_release_cache() returns the memory.
namespace {
struct HostPlan {
    int dev = -1, m = 0, n = 0, r = 0;
    unsigned f = 0;
    bool with_q = false;
    mpqr_handle* h = nullptr;
    float *dA = nullptr, *dQ = nullptr;
    bool busy = false;
};
std::mutex g_host_mu;
HostPlan g_host_plan;
void host_plan_free(HostPlan& p) {
    if (p.h) mpqr_destroy(p.h);  // (dA, dQ were allocated through the handle)
    p = HostPlan();
}
}  // namespace

int mpqr_debug_dump_trace(mpqr_handle* h);

int mpqr_release_cache(void) {
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        if (!g_host_plan.busy) host_plan_free(g_host_plan);
    }
    return mpqr_tsqr_release_cache();
}

int mpqr_block_qr_host(float* A_packed, float* Q, int m, int n, int r, unsigned flags) {
    if (!A_packed || m < 1 || n < 1 || r < 1) {
        set_error("mpqr_block_qr_host: bad arguments m=%d n=%d r=%d", m, n, r);
        return MPQR_EINVAL;
    }
    // (the streamed-input schedule of the 16-bit path needs every block's W for the catch-up of late chunks: MPQR_KEEP_WY
    //  storage; MPQR_NO_STREAM_IN=1 restores copy-then-factor)
    const bool want_stream_in = (flags & MPQR_PRECISION_MASK) != 0 && !getenv("MPQR_NO_STREAM_IN") && (long)m * n >= (1L << 24);
    const unsigned f = (flags & MPQR_PRECISION_MASK) | ((Q || want_stream_in) ? MPQR_KEEP_WY : 0u);
    const bool htrace = getenv("MPQR_HOST_TRACE") != nullptr;  // phase times of this call on stderr
    const bool no_cache = getenv("MPQR_NO_HOST_CACHE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_created = 0, t_h2d = 0, t_issued = 0;
    int dev = 0;
    MPQR_CUDA(cudaGetDevice(&dev));
    const long lda = round_up(n, 8), ldq = round_up(m, 8);
    HostPlan P;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        HostPlan& c = g_host_plan;
        if (!no_cache && !c.busy && c.h && c.dev == dev && c.m == m && c.n == n && c.r == r && c.f == f && c.with_q == (Q != nullptr)) {
            c.busy = true;
            P = c;
        } else if (!c.busy && c.h) {
            host_plan_free(c);  // another shape: one plan at a time
        }
    }
    const bool reused = P.h != nullptr;
    int rc = MPQR_OK;
    if (!reused) {
        P.dev = dev; P.m = m; P.n = n; P.r = r; P.f = f; P.with_q = Q != nullptr;
        rc = mpqr_create(&P.h, m, n, r, 0, f);
        if (rc == MPQR_OK) rc = dev_alloc(P.h, (void**)&P.dA, (size_t)(m + 1) * lda * sizeof(float));
        if (rc == MPQR_OK && Q) rc = dev_alloc(P.h, (void**)&P.dQ, (size_t)m * ldq * sizeof(float));
        if (rc != MPQR_OK) {
            if (P.h) mpqr_destroy(P.h);
            return rc;
        }
    }
    mpqr_handle* h = P.h;
    float *dA = P.dA, *dQ = P.dQ;
    t_created = now();
    do {
        cudaError_t e = cudaSuccess;
        // mixed path: finished column blocks go back to the host while later blocks are still being factored
        // (only for page-locked host buffers: an "async" copy to pageable memory blocks the issuing thread)
        cudaPointerAttributes pa{};
        const bool pinned = cudaPointerGetAttributes(&pa, A_packed) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        const bool pipelined = pinned && (f & MPQR_PRECISION_MASK) != 0;
        // Streamed input: column chunks of whole outer blocks (1, 1, 2, then 4 blocks each) are copied on their own
        // stream, each followed by its 16-bit shadow conversion and an event; the look-ahead driver starts on block 0 as
        // soon as it has landed and admits the later chunks as they arrive (factor_16).  PCIe is the limit either way
        // (4.3 GB at ~55 GB/s = 79 ms at 32768^2), but it now runs next to the factorisation instead of in front of it.
        h->arr.on = false;
        const int nblk_outer = ceil_div(h->kmax, h->nb);
        if (pipelined && want_stream_in && h->ov.on && nblk_outer >= 4 && h->keep_wy) {
            auto& ar = h->arr;
            if (!ar.stream && cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); rc = MPQR_ECUDA; break; }
            ar.c0.clear(); ar.c1.clear(); ar.t_ms.clear();
            constexpr double gbs = 52.0;  // expected host-to-device rate ([B200] 51-52 GB/s measured for pinned memory)
            double t = 0.05;
            for (int blk = 0, k = 0; blk * h->nb < n; ++k) {
                const int nblocks = k == 0 ? 1 : (k == 1 ? 3 : (k == 2 ? 4 : 8));   // few, growing chunks: every distant chunk has its own stream
                // ([B200, r2v] finer schedules lose: 1,3,4,4,4,8,4,2,2 blocks 227 ms, 1,3,4,8,8,4,2,2 234 ms, against 167 ms: every
                //  distant chunk adds a far update per outer block to the update partition)
                const int a0 = blk * h->nb;
                int a1 = (blk + nblocks) * h->nb;
                if (a1 > n || n - a1 < h->nb) a1 = n;   // a short tail joins the last chunk
                ar.c0.push_back(a0); ar.c1.push_back(a1);
                t += (double)(m + 1) * (a1 - a0) * 4.0 / (gbs * 1e6);
                ar.t_ms.push_back(t);
                blk = (a1 + h->nb - 1) / h->nb;
                if (a1 >= n) break;
            }
            while (ar.ev.size() < ar.c0.size()) {
                cudaEvent_t ev;
                if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { rc = MPQR_ECUDA; break; }
                ar.ev.push_back(ev);
            }
            if (rc != MPQR_OK) { set_error("event creation failed"); break; }
            std::vector<cudaEvent_t> tev;   // MPQR_HOST_TRACE: measured arrival times
            if (htrace) {
                tev.resize(ar.c0.size() + 1);
                for (auto& t2 : tev) cudaEventCreate(&t2);
                cudaEventRecord(tev[0], ar.stream);
            }
            h->arr_trace = tev;
            for (size_t k = 0; k < ar.c0.size() && e == cudaSuccess; ++k) {
                const int a0 = ar.c0[k], wdt = ar.c1[k] - ar.c0[k];
                e = cudaMemcpy2DAsync(dA + a0, lda * sizeof(float), A_packed + a0, (size_t)n * sizeof(float), (size_t)wdt * sizeof(float),
                                      m + 1, cudaMemcpyHostToDevice, ar.stream);
                if (e == cudaSuccess && convert_f32_to_16(dA + a0, lda, at16(h->Ah, h->ldh, 0, a0), h->ldh, m, wdt, h->prec == 2, ar.stream) != MPQR_OK) e = cudaErrorUnknown;
                if (e == cudaSuccess) e = cudaEventRecord(ar.ev[k], ar.stream);
                if (htrace) cudaEventRecord(h->arr_trace[k + 1], ar.stream);
            }
            if (e != cudaSuccess) { set_error("streamed H2D failed: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; break; }
            ar.on = true;
        } else {
            e = cudaMemcpy2D(dA, lda * sizeof(float), A_packed, (size_t)n * sizeof(float), (size_t)n * sizeof(float),
                             m + 1, cudaMemcpyHostToDevice);
        }
        t_h2d = now();
        if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; break; }
        h->sink_host = nullptr;
        if (pipelined) {
            if (!h->sink_stream && cudaStreamCreateWithFlags(&h->sink_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); rc = MPQR_ECUDA; break; }
            h->sink_host = A_packed;
            h->sink_pitch = (size_t)n * sizeof(float);
        }
        if ((rc = mpqr_factor_device(h, dA, lda, nullptr))) break;
        if (Q && (rc = mpqr_form_q_device(h, dQ, ldq, nullptr))) break;
        t_issued = now();
        if (pipelined) {
            e = cudaStreamSynchronize(h->sink_stream);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
        } else {
            e = cudaMemcpy2D(A_packed, (size_t)n * sizeof(float), dA, lda * sizeof(float), (size_t)n * sizeof(float), m + 1,
                             cudaMemcpyDeviceToHost);
        }
        if (e == cudaSuccess && Q)
            e = cudaMemcpy2D(Q, (size_t)m * sizeof(float), dQ, ldq * sizeof(float), (size_t)m * sizeof(float), m,
                             cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("D2H copy / kernel execution failed: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; break; }
    } while (0);
    h->sink_host = nullptr;
    h->arr.on = false;
    const double t_done = now();
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        HostPlan& c = g_host_plan;
        if (reused) {
            c.busy = false;
            if (rc != MPQR_OK) host_plan_free(c);  // do not keep a plan whose last run failed
        } else if (rc == MPQR_OK && !no_cache && !c.h) {
            c = P;
            c.busy = false;
        } else {
            mpqr_destroy(P.h);
        }
    }
    if (htrace && !h->arr_trace.empty()) {
        fprintf(stderr, "  chunk arrivals (ms after the first copy was issued; expected):");
        for (size_t k = 1; k < h->arr_trace.size(); ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, h->arr_trace[0], h->arr_trace[k]);
            fprintf(stderr, " [%d,%d) %.1f (%.1f)", h->arr.c0[k - 1], h->arr.c1[k - 1], t, h->arr.t_ms[k - 1]);
        }
        fprintf(stderr, "\n");
        for (auto t2 : h->arr_trace) cudaEventDestroy(t2);
        h->arr_trace.clear();
        if (h->ov.trace) mpqr_debug_dump_trace(h);
    }
    if (htrace) {
        fprintf(stderr, "  host issue by category (ms): chain+side %.1f, finalize/G/T/W %.1f, in-block %.1f, accumulation %.1f, far (incl. distant) %.1f, sink %.1f, single-block panels %.1f\n",
                g_host_prof[0], g_host_prof[1], g_host_prof[2], g_host_prof[3], g_host_prof[4], g_host_prof[6], g_host_prof[7]);
        for (double& v : g_host_prof) v = 0;
    }
    if (htrace)
        fprintf(stderr, "mpqr_block_qr_host %dx%d: plan %.1f ms (%s), H2D %.1f, issue %.1f, wait+D2H %.1f, release %.1f, total %.1f\n",
                m, n, t_created - t_begin, reused ? "cached" : "created", t_h2d - t_created, t_issued - t_h2d, t_done - t_issued,
                now() - t_done, now() - t_begin);
    return rc;
}

int mpqr_panel_factor_device(float* dA, long lda, int m, int n, int lam, int pw, float* dY, float* dW, float* dT,
                             void* stream) {
    if (!dA || lda < n) {
        set_error("mpqr_panel_factor_device: bad arguments");
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    float *ws = nullptr, *scratch = nullptr, *pws = nullptr;
    MPQR_CUDA(cudaMalloc(&ws, panel_sync_ws_bytes()));
    MPQR_CUDA(cudaMemset(ws, 0, panel_sync_ws_bytes()));
    if (cudaMalloc(&pws, panel_ws_bytes(m - lam)) != cudaSuccess) {
        cudaFree(ws);
        set_error("panel workspace allocation failed");
        return MPQR_ENOMEM;
    }
    unsigned host_ctr = 0;
    long scratch_rows = 0;
    if (m - lam > 32768) {
        scratch_rows = m;
        if (cudaMalloc(&scratch, panel_scratch_bytes(m)) != cudaSuccess) {
            cudaFree(ws); cudaFree(pws);
            set_error("panel scratch allocation failed");
            return MPQR_ENOMEM;
        }
    }
    PanelArgs a{};
    a.A = dA; a.lda = lda; a.m = m; a.n = n; a.lam = lam; a.acol = lam; a.pw = pw; a.blk_row0 = lam;
    a.Y32 = dY; a.W32 = dW; a.ld32 = pw;
    if (dW && !dY) { set_error("mpqr_panel_factor_device: dW needs dY"); cudaFree(ws); cudaFree(scratch); cudaFree(pws); return MPQR_EINVAL; }
    a.T = dT; a.ldt = pw;
    a.sync_ws = ws; a.host_ctr = &host_ctr; a.scratch = scratch; a.scratch_rows = scratch_rows;
    a.ws = pws; a.ws_rows = m - lam;
    unsigned* cflags = nullptr;
    unsigned cctr = 0;
    cudaStream_t cside = nullptr;
    if (cudaMalloc(&cflags, 64) == cudaSuccess && cudaMemset(cflags, 0, 64) == cudaSuccess &&
        cudaStreamCreateWithFlags(&cside, cudaStreamNonBlocking) == cudaSuccess) {
        a.chain_side = cside; a.chain_flags = cflags; a.chain_ctr = &cctr;
    } else {
        cudaGetLastError();
    }
    unsigned clast = 0;
    a.chain_last_far = &clast;
    int rc = launch_panel(a, (cudaStream_t)stream, nullptr);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (cside) { cudaStreamSynchronize(cside); cudaStreamDestroy(cside); }
    cudaFree(cflags);
    cudaFree(ws);
    cudaFree(scratch);
    cudaFree(pws);
    if (rc == MPQR_OK && e != cudaSuccess) {
        set_error("panel kernel failed: %s", cudaGetErrorString(e));
        rc = MPQR_ECUDA;
    }
    return rc;
}

// MPQR_TRACE=1: per-interval timeline of the last look-ahead factorisation (tools/quick_time.py TRACE=1)
int mpqr_debug_dump_trace(mpqr_handle* h) {
    if (!h || !h->ov.on || !h->ov.trace) return MPQR_ESTATE;
    cudaDeviceSynchronize();
    auto& o = h->ov;
    const int nblk = (int)o.tr.size();
    fprintf(stderr, "# b  panelSMs  t0(bp)  bp_ms  | t0(fn)  fn_ms  fr_ms   (times relative to bp(0) start)\n");
    for (int b = 0; b < nblk; ++b) {
        float tb0 = 0, bp = 0, tf0 = -1, fn = -1, fr = -1;
        cudaEventElapsedTime(&tb0, o.tr[0].b0, o.tr[b].b0);
        cudaEventElapsedTime(&bp, o.tr[b].b0, o.tr[b].b1);
        if (b + 1 < nblk || h->n > h->kmax) {
            if (cudaEventElapsedTime(&tf0, o.tr[0].b0, o.tr[b].f0) != cudaSuccess) { cudaGetLastError(); tf0 = -1; }
            if (cudaEventElapsedTime(&fn, o.tr[b].f0, o.tr[b].f1) != cudaSuccess) { cudaGetLastError(); fn = -1; }
            if (cudaEventElapsedTime(&fr, o.tr[b].f1, o.tr[b].f2) != cudaSuccess) { cudaGetLastError(); fr = -1; }
        }
        fprintf(stderr, "%3d  %4d  %8.2f %7.2f | %8.2f %6.2f %7.2f\n", b, o.tr[b].psm, tb0, bp, tf0, fn, fr);
    }
    return MPQR_OK;
}

// Tuning/profiling hook (tools/panel_probe.py; not part of the public header): runs the panel
// path once with phase profiling enabled.  dDbg: 16 x int64 on the device.
int mpqr_debug_panel_probe(float* dA, long lda, int m, int n, int lam, int pw, int force_b, int force_cs, int force_rpt,
                           int want_wy, long long* dDbg, void* stream) {
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    float *Y = nullptr, *W = nullptr, *T = nullptr, *pws = nullptr;
    if (want_wy) {
        MPQR_CUDA(cudaMalloc(&Y, (size_t)(m - lam) * pw * 4));
        MPQR_CUDA(cudaMalloc(&W, (size_t)(m - lam) * pw * 4));
        MPQR_CUDA(cudaMalloc(&T, (size_t)pw * pw * 4));
    }
    MPQR_CUDA(cudaMalloc(&pws, panel_ws_bytes(m - lam)));
    PanelArgs a{};
    a.A = dA; a.lda = lda; a.m = m; a.n = n; a.lam = lam; a.acol = lam; a.pw = pw; a.blk_row0 = lam;
    a.Y32 = Y; a.W32 = W; a.ld32 = pw; a.T = T; a.ldt = pw;
    a.ws = pws; a.ws_rows = m - lam;
    a.force_b = force_b; a.force_cs = force_cs; a.force_rpt = force_rpt;
    int caps[3] = {0, 0, 0};
    a.dbg_caps = caps;
    // force_b == -1: probe the persistent chain (dDbg receives 8 globaltimer stamps per register block) instead of the
    // phase counters of the single-block kernels
    unsigned* cflags = nullptr;
    unsigned cctr = 0;
    cudaStream_t cside = nullptr;
    if (force_b == -1) {
        a.force_b = 0;
        a.chain_dbg = dDbg;
        MPQR_CUDA(cudaMalloc(&cflags, 64));
        MPQR_CUDA(cudaMemset(cflags, 0, 64));
        MPQR_CUDA(cudaStreamCreateWithFlags(&cside, cudaStreamNonBlocking));
        a.chain_side = cside; a.chain_flags = cflags; a.chain_ctr = &cctr;
    } else {
        a.dbg = dDbg;
    }
    int rc = launch_panel(a, (cudaStream_t)stream, nullptr);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (cside) { cudaStreamSynchronize(cside); cudaStreamDestroy(cside); }
    cudaFree(cflags);
    cudaFree(Y); cudaFree(W); cudaFree(T); cudaFree(pws);
    if (force_b == -1) return (rc == MPQR_OK && e != cudaSuccess) ? MPQR_ECUDA : rc;
    if (rc == MPQR_OK && e != cudaSuccess) { set_error("panel probe failed: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; }
    if (rc == MPQR_OK && dDbg) {
        long long c[3] = {caps[0], caps[1], caps[2]};
        cudaMemcpy(dDbg + 13, c, sizeof(c), cudaMemcpyHostToDevice);
    }
    return rc;
}

int mpqr_gemm_tn_device(const void* dX, long ldx, const void* dZ, long ldz, float* dS, long lds, int M, int N, int K,
                        int bf16, void* stream) {
    if (!dX || !dZ || !dS || K < 1) { set_error("mpqr_gemm_tn_device: bad arguments"); return MPQR_EINVAL; }
    return tc_gemm_tn(dX, ldx, dZ, ldz, dS, lds, M, N, K, bf16, 0, (cudaStream_t)stream, nullptr);
}

int mpqr_gemm_nn_device(const void* dX, long ldx, const void* dS16, long lds16, float* dC, long ldc, void* dC16,
                        long ldc16, int M, int N, int K, int bf16, void* stream) {
    if (!dX || !dS16 || !dC) { set_error("mpqr_gemm_nn_device: bad arguments"); return MPQR_EINVAL; }
    return tc_gemm_nn(dX, ldx, dS16, lds16, dC, ldc, dC16, ldc16, M, N, K, bf16, 0, (cudaStream_t)stream, nullptr);
}

int mpqr_fill_uniform_device(float* dA, long lda, long n_total, long row0, long rows, long col0, long cols,
                             uint64_t seed, void* stream) {
    if (!dA) { set_error("mpqr_fill_uniform_device: null pointer"); return MPQR_EINVAL; }
    return fill_uniform(dA, lda, n_total, row0, rows, col0, cols, seed, (cudaStream_t)stream);
}

}  // extern "C"
