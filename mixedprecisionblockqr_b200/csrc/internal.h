// internal.h — handle layout and helpers shared by api.cu and mg.cu (not installed).
#pragma once
#include <vector>

#include "common.cuh"

// ------------------------------------------------------------------------------ the handle
struct mpqr_handle {
    int m = 0, n = 0, r = 0, nb = 0, kmax = 0;
    unsigned flags = 0;
    int prec = 0;  // 0 fp32, 1 fp16, 2 bf16
    bool keep_wy = false;
    int npanels = 0;
    bool factored = false;
    long launches = 0;

    // common
    float* sync_ws = nullptr;
    unsigned sync_ctr = 0;  // host mirror of the panel barrier counter
    float* scratch = nullptr;
    long scratch_rows = 0;
    float* panel_ws = nullptr;  // panel.cu workspace (in-panel blocks, Gram, T)
    long panel_ws_rows = 0;
    // persistent panel chain (panel.cu): device flags [done, far], their host mirror, the default side stream
    unsigned* chain_flags = nullptr;
    unsigned chain_ctr = 0;
    unsigned chain_last_far = 0;   // last value posted to chain_flags[1]
    unsigned chain_panels = 0;     // chain panels issued so far (picks the Y / T workspace buffer)
    cudaStream_t chain_side = nullptr;
    bool no_chain = false;         // MPQR_STREAM_ORDERED: never launch the persistent (flag-waiting) panel kernel
    float* T = nullptr;    // npanels * r * r
    float* S32 = nullptr;  // sk x lds32
    long lds32 = 0;
    int sk = 0;

    // FP32 path: compact Y/W.  keep_wy: m x ld32 full arrays, else m x r panel buffers
    float* Y32 = nullptr;
    float* W32 = nullptr;
    long ld32 = 0;

    // 16-bit path
    void* Ah = nullptr;  // m x ldh shadow of A (operands); factored columns hold Y
    long ldh = 0;
    void* W16 = nullptr;  // keep_wy: m x ldw16 (all blocks) else m x nb (current block)
    long ldw16 = 0;
    float* Wblk32 = nullptr;  // m x ldwb FP32 W of the current outer block
    long ldwb = 0;
    void* S16 = nullptr;      // sk x lds16
    long lds16 = 0;
    void* Qh = nullptr;  // m x ldqh shadow of Q (form_q)
    long ldqh = 0;

    // look-ahead driver (api.cu): two green-context SM partitions with one stream each
    // Several (panel, update) SM partitions are kept ready; the driver picks one per outer block.
    struct Overlap {
        bool on = false;
        struct Pair {
            void* gP = nullptr;  // CUgreenCtx: panel partition
            void* gU = nullptr;  // CUgreenCtx: update partition
            cudaStream_t sP = nullptr, sP2 = nullptr, sP3 = nullptr, sU = nullptr;  // sP2 (in-block rest), sP3 (chain side): more streams of the panel partition
            int nsmP = 0, nsmU = 0;
        };
        std::vector<Pair> pairs;
        cudaStream_t sF = nullptr, sF2 = nullptr, sF3 = nullptr;  // whole-device streams (intervals that are not worth splitting)
        std::vector<cudaEvent_t> ev_rest;  // BlockCtx::rest_ev
        int nsm_full = 0;
        std::vector<cudaEvent_t> ev_bp, ev_fn, ev_fr;
        cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_accdone = nullptr;
        std::vector<cudaEvent_t> ev_acc;  // per panel of an outer block: in-block update done -> WY accumulation may start
        // MPQR_TRACE=1: timed events around block_phase / far_next / far_rest of every interval
        bool trace = false;
        struct Tr { cudaEvent_t b0, b1, f0, f1, f2; int psm; };
        std::vector<Tr> tr;
    } ov;
    float* S32r = nullptr;  // scratch of the in-block look-ahead stream (r x lds32)
    void* S16r = nullptr;
    float* S32u = nullptr;  // scratch of the update stream (same shape as S32 / S16)
    void* S16u = nullptr;
    void* W16b = nullptr;   // second W buffer (blocks alternate) when the look-ahead driver is on

    // host sink (mpqr_block_qr_host): finished column blocks are copied back while later blocks compute
    float* sink_host = nullptr;
    size_t sink_pitch = 0;  // bytes
    cudaStream_t sink_stream = nullptr;
    std::vector<cudaEvent_t> sink_ev;

    // host source (mpqr_block_qr_host): the input arrives in column chunks while the first blocks are already being
    // factored; far updates only cover the chunks that have been admitted, a late chunk catches up on admission
    struct Arrival {
        bool on = false;
        std::vector<int> c0, c1;          // column range of each chunk (whole outer blocks, in order)
        std::vector<cudaEvent_t> ev;      // chunk is in HBM and its 16-bit shadow is written
        std::vector<double> t_ms;         // expected arrival time after the start of the call
        cudaStream_t stream = nullptr;    // copy + shadow conversion
        // per chunk (from chunk 1 on): own stream in the update partition, "blocks applied so far" event, GEMM scratch
        std::vector<cudaStream_t> cs;
        std::vector<cudaEvent_t> cev;
        std::vector<char> cev_set;
        std::vector<float*> cS32;
        std::vector<void*> cS16;
    } arr;
    std::vector<cudaEvent_t> arr_trace;   // MPQR_HOST_TRACE=1: timed events behind every chunk

    // multi-GPU (mg.cu)
    void* mg = nullptr;

    // per-class event profiling
    bool prof = false;
    struct ProfRec { int cls; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_flops[MPQR_NUM_KERNEL_CLASSES] = {};
    double prof_bytes[MPQR_NUM_KERNEL_CLASSES] = {};
    mpqr::ProfHook hook{};  // filled by mpqr_set_profiling

    std::vector<void*> allocs;
};


namespace mpqr {

int dev_alloc(mpqr_handle* h, void** p, size_t bytes);
void mg_destroy(void* state);

// RAII event pair around one launch (only when profiling is on)
struct ProfScope {
    mpqr_handle* h;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    ProfScope(mpqr_handle* h_, int cls, cudaStream_t st_, double flops, double bytes) : h(h_), st(st_) {
        if (!h->prof) return;
        cudaEvent_t e0;
        auto get = [&](cudaEvent_t* e) {
            if (!h->prof_pool.empty()) { *e = h->prof_pool.back(); h->prof_pool.pop_back(); }
            else cudaEventCreate(e);
        };
        get(&e0);
        get(&e1);
        cudaEventRecord(e0, st);
        h->prof_recs.push_back({cls, e0, e1});
        h->prof_flops[cls] += flops;
        h->prof_bytes[cls] += bytes;
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, st);
    }
};
#define PROF(cls, flops, bytes, call)                         \
    do {                                                      \
        ProfScope ps__(h, cls, st, flops, bytes);             \
        MPQR_TRY(call);                                       \
    } while (0)

// ProfHook implementation on the handle (used by launch_panel for its sub-classes)
inline void hook_begin(void* ctx, int cls, cudaStream_t st, double flops, double bytes) {
    mpqr_handle* h = (mpqr_handle*)ctx;
    if (!h->prof) return;
    cudaEvent_t e0, e1;
    auto get = [&](cudaEvent_t* e) {
        if (!h->prof_pool.empty()) { *e = h->prof_pool.back(); h->prof_pool.pop_back(); }
        else cudaEventCreate(e);
    };
    get(&e0);
    get(&e1);
    cudaEventRecord(e0, st);
    h->prof_recs.push_back({cls, e0, e1});
    h->prof_flops[cls] += flops;
    h->prof_bytes[cls] += bytes;
}
inline void hook_end(void* ctx, cudaStream_t st) {
    mpqr_handle* h = (mpqr_handle*)ctx;
    if (!h->prof || h->prof_recs.empty()) return;
    cudaEventRecord(h->prof_recs.back().e1, st);
}

inline double tn_bytes(double M, double N, double K) { return 2.0 * K * (M + N) + 4.0 * M * N; }
inline double nn_bytes(double M, double N, double K) { return 10.0 * M * N + 2.0 * K * (M + N); }


inline char* at16(void* base, long ld, long row, long col) { return (char*)base + ((size_t)row * ld + col) * 2; }

// One outer block [c0, c1) of the two-level driver.  Row indices are global; column indices
// refer to the arrays given here (single GPU: global columns; multi GPU: local columns).
struct BlockCtx {
    float* A;      // packed FP32 master
    long lda;
    int acol0;     // column of A (and of the shadow) holding global column c0
    void* Ah;      // 16-bit shadow of A
    long ldh;
    void* Y16;     // Y of the block: element (row c0, block column 0)
    long ldy;
    void* W16;     // W of the block: element (row c0, block column 0)
    long ldw;
    float* S32;    // GEMM scratch of the issuing stream (null: the handle's)
    void* S16;
    // Optional: run the WY accumulation of the block (W_p <- W_p - W_prev (Y_prev^T W_p), only needed by
    // the FAR update) on another stream / SM partition, off the panel chain.  The look-ahead driver points
    // it at the update partition, which idles while the schedule is panel-bound.
    cudaStream_t acc_stream;   // null: accumulate in line
    int acc_sms;               // SM budget of acc_stream (0 = whole device)
    float* acc_S32;            // its GEMM scratch
    void* acc_S16;
    cudaEvent_t* acc_ev;       // >= (c1 - c0) / r events
    // Optional: in-block look-ahead.  The in-block update of panel p is split: the NEXT panel's columns on the
    // panel stream, the other in-block columns on `rest_stream` (a second stream of the same SM partition), so
    // that the next panel's register-block kernels (16 SMs) run while the rest of the block is updated.
    cudaStream_t rest_stream;  // null: one in-block update on the panel stream
    float* rest_S32;           // its GEMM scratch (r x lds32)
    void* rest_S16;
    cudaEvent_t* rest_ev;      // 2 events per panel: [2p] panel factored, [2p+1] rest of panel p's in-block update done
    int rest_sms;              // SM budget of the GEMMs issued on rest_stream (0: the caller's budget).  They run NEXT TO the
                               // next panel's cluster and its side updates: a persistent GEMM grid over the whole partition
                               // kept the 16-CTA cluster waiting for ~47 us per panel at 32768 rows (profiles/r2_timeline_c4.txt)
    cudaStream_t chain_side;   // side stream of the persistent panel chain (null: the handle's own)
};
// panels + in-block updates + WY accumulation of block [c0, c1); `ncols_in` = columns of A
// (starting at acol0) that belong to the block's own panel region (= c1 - c0)
int block_phase(mpqr_handle* h, const BlockCtx& c, int c0, int c1, int n_end_is_matrix_end, cudaStream_t st);
// A[c0:, afar : afar+nfar) -= Y (W^T A[c0:, afar : afar+nfar))
int far_update(mpqr_handle* h, const BlockCtx& c, int c0, int c1, int afar, int nfar, cudaStream_t st);

}  // namespace mpqr
