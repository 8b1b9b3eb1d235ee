// common.cuh — shared declarations of libmpqr (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mpqr.h"

namespace mpqr {

void set_error(const char* fmt, ...);

#define MPQR_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ::mpqr::set_error("CUDA error at %s:%d code=%d(%s) \"%s\"", __FILE__, __LINE__,   \
                              (int)e__, cudaGetErrorName(e__), #expr);                        \
            return MPQR_ECUDA;                                                                \
        }                                                                                     \
    } while (0)

#define MPQR_TRY(expr)               \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != MPQR_OK) return rc__; \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long ceil_divl(long a, long b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

struct DeviceInfo {
    int num_sms;
    int max_smem_optin;
    int coop;
};
int get_device_info(DeviceInfo* out);

// SM budget of the stream the host is currently issuing into (green-context partitions of the
// look-ahead driver, api.cu).  0 = the whole device.  Persistent-kernel grids and split-K factors
// are sized from sm_count(), not from the device attribute.
extern thread_local int g_sm_budget;
struct SmBudget {
    int prev;
    explicit SmBudget(int n) : prev(g_sm_budget) { g_sm_budget = n; }
    ~SmBudget() { g_sm_budget = prev; }
};
inline int sm_count(const DeviceInfo& di) { return g_sm_budget > 0 ? g_sm_budget : di.num_sms; }

// Host-side issue profile (MPQR_HOST_TRACE=1): wall time the issuing thread spends per category of calls.
//   0 chain kernel + side-stream items   1 finalize + Gram / T / W   2 in-block GEMMs   3 WY accumulation
//   4 far updates   5 distant-chunk updates (streamed input)   6 D2H sink   7 panel (single block / classic flow)
extern thread_local double g_host_prof[8];
extern bool g_host_prof_on;
struct HostProfScope {
    int cat;
    double t0;
    explicit HostProfScope(int c);
    ~HostProfScope();
};

// cudaFuncSetAttribute(func, attr, value) once per (function, attribute, device): function attributes are per device,
// and the library is used on several devices from one process (host plan cache, TSQR lanes, multi-GPU tests).
int func_attr_once(const void* func, cudaFuncAttribute attr, int value);

// Programmatic dependent launch (PDL): every kernel of the factorisation chain lets its successor
// become resident at once (launch_dependents) and waits for its predecessor's completion and
// memory flush (wait) only after its own set-up, so launch latency and prologues overlap the
// predecessor's tail.  Host side: launch with pdl_attr() through cudaLaunchKernelEx.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
inline cudaLaunchAttribute pdl_attr() {
    static const int off = getenv("MPQR_NO_PDL") ? 1 : 0;  // debugging aid: plain stream order
    cudaLaunchAttribute a{};
    a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a.val.programmaticStreamSerializationAllowed = off ? 0 : 1;
    return a;
}

// ------------------------------------------------------------------ panel factorisation
constexpr int kPanelMaxWidth = 128;

// Sync workspace of the panel kernel: per-CTA slots (double buffered) | Gram accumulator |
// barrier counter.  Zeroed ONCE when allocated; the counter only grows (host_ctr mirrors it).
size_t panel_sync_ws_bytes();
// Scratch for the non-shared-memory-resident variant: rows x 128 floats.
size_t panel_scratch_bytes(int max_rows);

// optional per-launch event hook (internal.h implements it on the handle's profiler)
struct ProfHook {
    void* ctx;
    void (*begin)(void* ctx, int cls, cudaStream_t st, double flops, double bytes);
    void (*end)(void* ctx, cudaStream_t st);
};

struct PanelArgs {
    float* A;      // packed master, (m+1) x lda
    long lda;
    int m, n;
    int lam;       // first ROW of the panel = global index of its first column
    int acol;      // column of A holding the panel's first column (== lam on a single GPU)
    int pw;        // panel width (<= 128), lam + pw <= n
    int blk_row0;  // first row of the enclosing outer block (<= lam): Y/W outputs are zero-filled
                   // for rows [blk_row0, lam)
    // Compact outputs at UNSHIFTED positions; each pointer addresses element
    // (row = blk_row0, col = first panel column) of its array.  Any may be null.
    float* Y32;
    float* W32;
    long ld32;
    void* Y16;     // __half or __nv_bfloat16
    void* W16;
    long ldy16, ldw16;
    int bf16;
    float* T;      // pw x pw (ldt), upper triangular, Q_p = I - Y T Y^T
    int ldt;
    float* sync_ws;
    unsigned* host_ctr; // host-side mirror of the barrier counter in sync_ws (running total)
    unsigned ctr_base;  // filled in by launch_panel
    float* scratch;     // may be null if the panel fits in shared memory
    long scratch_rows;  // capacity of scratch in rows
    long long* dbg;     // optional device buffer (16 x int64) for phase profiling, else null
    int rows_hint;      // > 0: override the rows-per-CTA heuristic (tuning)
    int force_cs;       // > 0: cap the cluster size (tuning / tests)
    int* dbg_caps;      // optional HOST array[3]: max cluster size, chosen cluster size, rows per thread
    float* ws;          // panel workspace (panel_ws_bytes), needed when pw > one register block
    long ws_rows;       // rows the workspace was sized for
    int force_b;        // 16 / 32: override the register-block width (tuning / tests)
    int force_rpt;      // > 0: override the rows per thread (tuning / tests)
    const ProfHook* prof;  // optional: sub-classes 4 (block kernels), 5 (in-panel S), 6 (Gram/T/W), 7 (in-panel U)
    // Persistent panel chain (panel_chain_kernel): one cluster launch per panel on `stream`; the updates of the rest of
    // the panel run on `chain_side`, ordered against the running kernel through two device flags.
    cudaStream_t chain_side;  // null: the chain flow is not used
    unsigned* chain_flags;    // device: [0] block done (kernel -> side stream), [1] far update done (side stream -> kernel)
    unsigned* chain_ctr;      // host mirror: the flags only grow
    long long* chain_dbg;     // optional device buffer, 8 x int64 per register block: globaltimer stamps of the cluster
    int chain_buf;            // which of the two Y / T workspace buffers this panel uses (consecutive panels alternate)
    unsigned* chain_last_far; // host: last value posted to chain_flags[1] so far (0 = none); updated by launch_panel
    // Merged Gram / next-panel product (mixed path; the 16-bit Y lives in the dead columns of the operand shadow, so the
    // gs_ncols columns right of the panel in the SAME array are the shadow of the next panel's columns): ONE TN GEMM gives
    // [G | Sy] = Y^T [Y | A_next], the T kernel also writes S = T^T Sy (16-bit, gs_S16: gs_ncols columns, ld gs_lds16),
    // which is what the in-block update A_next -= Y S needs -- W = Y T is not on the way to the next panel any more.
    int gs_ncols;             // 0: plain Gram
    void* gs_S16;
    long gs_lds16;
    int defer_w;              // 1: launch_panel does not form W; the caller issues panel_form_w (any stream behind `stream`)
};
// W = Y T of a panel factored with defer_w (same arguments; mixed path)
int panel_form_w(const PanelArgs& a, cudaStream_t stream, long* launches);
// CUDA loads kernels lazily and a first-time load may need the device to drain: while a gate kernel / the cluster spins on
// a flag that a later launch has to satisfy, that load would never return.  Every kernel a factorisation can launch is
// therefore loaded (and given its attributes) before the first one is issued on a device (chain_preload, panel.cu).
int preload_tc_gemm();
int preload_simt_gemm();
int preload_util_kernels();
int chain_preload_all();
// true if launch_panel will take the persistent chain flow for this panel
bool panel_chain_ok(const PanelArgs& a);
size_t panel_ws_bytes(long max_rows);
int launch_panel(const PanelArgs& a, cudaStream_t stream, long* launches);

// ------------------------------------------------------------------ FP32 SIMT GEMMs
// S[M x N] = X^T Z ; X [K x M] (ldx), Z [K x N] (ldz).  S is overwritten.
int sgemm_tn(const float* X, long ldx, const float* Z, long ldz, float* S, long lds, int M, int N,
             int K, cudaStream_t stream, long* launches);
// C[M x N] -= X S ; X [M x K] (ldx), S [K x N] (lds).
int sgemm_nn_sub(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M,
                 int N, int K, cudaStream_t stream, long* launches);

// C[M x N] = X S (store; the TSQR thin-Q products)
int sgemm_nn_store(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K,
                   cudaStream_t stream);

// 16-bit-operand CUDA-core fallbacks (sub-blocks that are not 16-byte aligned, see gemm_tc.cu)
int simt16_gemm_tn(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, int M, int N, int K,
                   int bf16, cudaStream_t stream);
int simt16_gemm_nn(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16, long ldc16,
                   int M, int N, int K, int bf16, cudaStream_t stream, int store = 0);

// ------------------------------------------------------------------ tcgen05 GEMMs
// S[M x N] (fp32) = X^T Z with 16-bit X [K x M], Z [K x N]; split-K with TMA reduce-add when
// the tile count cannot fill the GPU (S is zeroed internally in that case).
int tc_gemm_tn(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, int M, int N,
               int K, int bf16, int pad_ok, cudaStream_t stream, long* launches);
// The same with a 16-bit copy of S for the following NN GEMM: without split-K the epilogue writes S16 directly
// (*wrote16 = 1, S untouched), else S is produced and the caller converts.
int tc_gemm_tn16(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, void* S16, long lds16, int* wrote16,
                 int M, int N, int K, int bf16, int pad_ok, cudaStream_t stream, long* launches);
// C[M x N] (fp32) -= X S16, X [M x K] 16-bit (K-major), S16 [K x N] 16-bit; optionally mirrors
// the updated C into C16 (16-bit shadow).
int tc_gemm_nn(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16,
               long ldc16, int M, int N, int K, int bf16, int pad_ok, cudaStream_t stream, long* launches);

// C[M x N] (fp32) = X S16 (store), same operands; C16 optional 16-bit copy.
int tc_gemm_nn_store(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16,
                     long ldc16, int M, int N, int K, int bf16, cudaStream_t stream, long* launches);

// ------------------------------------------------------------------ small utility kernels
int fill_uniform(float* A, long lda, long n_total, long row0, long rows, long col0, long cols,
                 uint64_t seed, cudaStream_t stream);
int convert_f32_to_16(const float* src, long lds, void* dst, long ldd, long rows, long cols, int bf16,
                      cudaStream_t stream);
int set_identity(float* Q, long ldq, int m, cudaStream_t stream);
int fill_zero_16(void* dst, long ldd, long rows, long cols, cudaStream_t stream);

}  // namespace mpqr
