// mg.cu — multi-GPU driver: 1-D column-block-cyclic over the GPUs of one NVSwitch box, one
// process per GPU (the reference is single-GPU, Cuda/qr.cu has no device selection).
//
//   * global outer block b (columns [b*nb, (b+1)*nb)) lives on rank b % P; every rank stores
//     all m(+1) rows of its own blocks, concatenated in block order (row-major, lda_local);
//   * block b is factored by its owner exactly like on one GPU (block_phase: panel kernel +
//     in-block tcgen05 updates + WY accumulation), producing the block's Y and W as contiguous
//     16-bit staging buffers;
//   * owner broadcasts Y|W (2 * D * nb * 2 bytes) with ncclBroadcast over NVLink; every rank
//     applies the far update to its own later columns (far_update: the same tcgen05 GEMM pair).
//
// NCCL is dlopen()ed so that libmpqr.so loads on machines without it (CPU symbol tests); the
// communicator is owned by the handle.  The unique id is exchanged out of band (bench.py and
// the tests use torch.distributed).
#include <dlfcn.h>
#include <string.h>

#include "internal.h"

using namespace mpqr;

namespace {

// minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, result code 0 = success
typedef struct ncclComm* ncclComm_t;
struct NcclUid { char internal[MPQR_NCCL_UID_BYTES]; };
typedef int (*fn_GetUniqueId)(NcclUid*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, NcclUid, int);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_Broadcast)(const void*, void*, size_t, int /*dtype*/, int /*root*/, ncclComm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t /*sendcount*/, int /*dtype*/, ncclComm_t, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);
constexpr int kNcclChar = 0;

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_Broadcast Broadcast = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
};

int load_nccl(NcclApi** out) {
    static NcclApi api;
    if (!api.lib) {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            set_error("cannot dlopen libnccl.so.2: %s", dlerror());
            return MPQR_ENCCL;
        }
        api.GetUniqueId = (fn_GetUniqueId)dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (fn_CommInitRank)dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (fn_CommDestroy)dlsym(api.lib, "ncclCommDestroy");
        api.Broadcast = (fn_Broadcast)dlsym(api.lib, "ncclBroadcast");
        api.AllGather = (fn_AllGather)dlsym(api.lib, "ncclAllGather");
        api.GetErrorString = (fn_GetErrorString)dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllGather) {
            set_error("libnccl lacks a required symbol");
            api.lib = nullptr;
            return MPQR_ENCCL;
        }
    }
    *out = &api;
    return MPQR_OK;
}

struct MgState {
    NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int nloc = 0;        // local columns
    void* YW16[2] = {nullptr, nullptr};  // staging (blocks alternate): Y block then W block, each m x ldw 16-bit
    long ldw = 0;
    // look-ahead schedule: panel chain / far updates / NCCL on their own streams (green-context partitions when the
    // handle has them), ordered by events
    cudaStream_t s_comm = nullptr, s_panel = nullptr, s_upd = nullptr;
    bool own_streams = false;
    std::vector<cudaEvent_t> ev_bp, ev_bc, ev_far, ev_next;
    cudaEvent_t ev_start = nullptr;
    void* Ah = nullptr;  // local shadow, m x ldh
    long ldh = 0;
    // TSQR scratch (grow-only, freed with the handle): R_p | R stack | Q of the stack | thin Q_p
    float* tq[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t tq_bytes[4] = {0, 0, 0, 0};
    int tq_get(int i, size_t bytes, float** out) {
        if (tq_bytes[i] < bytes) {
            cudaFree(tq[i]);
            tq[i] = nullptr;
            tq_bytes[i] = 0;
            if (cudaMalloc(&tq[i], bytes) != cudaSuccess) {
                cudaGetLastError();
                set_error("mpqr_mg_tsqr_device: device allocation of %zu bytes failed", bytes);
                return MPQR_ENOMEM;
            }
            tq_bytes[i] = bytes;
        }
        *out = tq[i];
        return MPQR_OK;
    }
};

#define MPQR_NCCL(api, expr)                                                                     \
    do {                                                                                         \
        int r__ = (expr);                                                                        \
        if (r__ != 0) {                                                                          \
            set_error("NCCL error %d (%s) at %s:%d", r__,                                        \
                      (api)->GetErrorString ? (api)->GetErrorString(r__) : "?", __FILE__, __LINE__); \
            return MPQR_ENCCL;                                                                   \
        }                                                                                        \
    } while (0)

}  // namespace

extern "C" {

// ---- pure host layout helpers (no CUDA needed; tests/test_mg_layout.py uses them on CPU)
// number of columns of an n-column matrix owned by `rank` with block width nb
int mpqr_mg_layout_local_cols(int n, int nb, int rank, int nranks) {
    if (n < 0 || nb < 1 || nranks < 1 || rank < 0 || rank >= nranks) return MPQR_EINVAL;
    int nblk = (n + nb - 1) / nb, cols = 0;
    for (int b = rank; b < nblk; b += nranks) cols += (b * nb + nb <= n) ? nb : n - b * nb;
    return cols;
}
// global column of local column j on `rank` (or -1)
int mpqr_mg_layout_global_col(int n, int nb, int rank, int nranks, int local_col) {
    if (nb < 1 || nranks < 1 || local_col < 0) return -1;
    int lb = local_col / nb, off = local_col % nb;
    int g = (lb * nranks + rank) * nb + off;
    return g < n ? g : -1;
}

int mpqr_mg_get_unique_id(void* uid_out) {
    NcclApi* api;
    MPQR_TRY(load_nccl(&api));
    NcclUid id;
    MPQR_NCCL(api, api->GetUniqueId(&id));
    memcpy(uid_out, &id, sizeof(id));
    return MPQR_OK;
}

int mpqr_mg_create(mpqr_handle** out, int m, int n, int r, int nb, unsigned flags, int rank, int nranks,
                   const void* uid) {
    if (!out || !uid || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("mpqr_mg_create: bad arguments");
        return MPQR_EINVAL;
    }
    if ((flags & MPQR_PRECISION_MASK) == 0) flags |= MPQR_FP16;  // the multi-GPU path is tensor-core only
    if (flags & MPQR_KEEP_WY) {
        set_error("mpqr_mg_create: MPQR_KEEP_WY / explicit Q is not available on the multi-GPU path yet");
        return MPQR_EINVAL;
    }
    mpqr_handle* h = nullptr;
    // The base handle provides T, the panel workspaces, S32/S16, Wblk32 and (for >= 4 outer blocks) the green-context
    // partitions; its full-matrix shadow / W16 buffers stay unused on this path (memory is not the constraint here:
    // 180 GB per GPU), the local shadow and the Y|W staging buffers are added below.
    int nb_eff = nb > 0 ? nb : 1024;
    MPQR_TRY(mpqr_create(&h, m, n, r, nb_eff, flags));
    MgState* g = new MgState();
    h->mg = g;
    g->rank = rank;
    g->nranks = nranks;
    g->nloc = mpqr_mg_layout_local_cols(n, h->nb, rank, nranks);
    int rc = MPQR_OK;
    do {
        if ((rc = load_nccl(&g->api))) break;
        NcclUid id;
        memcpy(&id, uid, sizeof(id));
        int nr = g->api->CommInitRank(&g->comm, nranks, id, rank);
        if (nr != 0) {
            set_error("ncclCommInitRank failed: %d", nr);
            rc = MPQR_ENCCL;
            break;
        }
        g->ldw = round_up(h->nb, 8);
        if ((rc = dev_alloc(h, &g->YW16[0], (size_t)2 * m * g->ldw * 2))) break;
        if ((rc = dev_alloc(h, &g->YW16[1], (size_t)2 * m * g->ldw * 2))) break;
        g->ldh = round_up(g->nloc > 0 ? g->nloc : 8, 8);
        if ((rc = dev_alloc(h, &g->Ah, (size_t)m * g->ldh * 2))) break;
        // GEMM scratch of the update / in-block-rest streams (the base handle only has them when its own look-ahead is on)
        if (!h->S32u && (rc = dev_alloc(h, (void**)&h->S32u, (size_t)h->sk * h->lds32 * sizeof(float)))) break;
        if (!h->S16u && (rc = dev_alloc(h, &h->S16u, (size_t)h->sk * h->lds16 * 2))) break;
        if (!h->S32r && (rc = dev_alloc(h, (void**)&h->S32r, (size_t)h->sk * h->lds32 * sizeof(float)))) break;
        if (!h->S16r && (rc = dev_alloc(h, &h->S16r, (size_t)h->sk * h->lds16 * 2))) break;
        if (cudaStreamCreateWithFlags(&g->s_comm, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); rc = MPQR_ECUDA; break; }
        if (!h->ov.on || h->ov.pairs.empty()) {
            if (cudaStreamCreateWithFlags(&g->s_panel, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&g->s_upd, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); rc = MPQR_ECUDA; break; }
            g->own_streams = true;
        }
        const int nblk = ceil_div(h->kmax, h->nb);
        g->ev_bp.resize(nblk); g->ev_bc.resize(nblk); g->ev_far.resize(nblk); g->ev_next.resize(nblk);
        for (int i = 0; i < nblk; ++i) {
            cudaEventCreateWithFlags(&g->ev_bp[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&g->ev_bc[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&g->ev_far[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&g->ev_next[i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&g->ev_start, cudaEventDisableTiming);
    } while (0);
    if (rc != MPQR_OK) {
        mpqr_destroy(h);
        return rc;
    }
    *out = h;
    return MPQR_OK;
}

int mpqr_mg_local_cols(const mpqr_handle* h) {
    if (!h || !h->mg) return MPQR_EINVAL;
    return ((MgState*)h->mg)->nloc;
}

int mpqr_mg_global_col(const mpqr_handle* h, int local_col) {
    if (!h || !h->mg) return MPQR_EINVAL;
    MgState* g = (MgState*)h->mg;
    return mpqr_mg_layout_global_col(h->n, h->nb, g->rank, g->nranks, local_col);
}

int mpqr_mg_factor_device(mpqr_handle* h, float* dA, long lda, void* stream) {
    if (!h || !h->mg || !dA) {
        set_error("mpqr_mg_factor_device: bad arguments");
        return MPQR_EINVAL;
    }
    MgState* g = (MgState*)h->mg;
    cudaStream_t st = (cudaStream_t)stream;
    const int m = h->m, n = h->n, nb = h->nb, P = g->nranks;
    const int bf = h->prec == 2;
    if (lda < g->nloc || (lda & 3) || ((uintptr_t)dA & 15)) {
        set_error("mpqr_mg_factor_device: needs lda_local >= local cols, lda %% 4 == 0, 16-byte aligned dA");
        return MPQR_EINVAL;
    }
    h->launches = 0;
    if (g->nloc > 0) {
        PROF(3, 0, 6.0 * m * g->nloc, convert_f32_to_16(dA, lda, g->Ah, g->ldh, m, g->nloc, bf, st));
        h->launches += 1;
    }
    // Look-ahead schedule: block b+1's owner updates that block's columns first (far_next on the panel
    // stream), factors it and broadcasts it on the comm stream WHILE every rank still applies block b to the rest of
    // its columns (far_rest on the update stream).  Y|W staging is double buffered.  Dependencies:
    //   bp(b)    : far_next(b-1) [same stream]; staging[b&1] free = far_rest(b-2), far_next(b-2) done
    //   bcast(b) : bp(b) on the owner; staging[b&1] free on the receivers
    //   far_*(b) : bcast(b); far_next(b) additionally far_rest(b-1) (same columns, other stream)
    const int nblk = ceil_div(h->kmax, nb);
    // streams: a (panel, update) green-context pair of the base handle when it has one, else plain streams
    cudaStream_t s_panel = g->s_panel, s_upd = g->s_upd, s_rest = nullptr, s_side = nullptr;
    int nsm_p = 0, nsm_u = 0;
    if (!g->own_streams) {
        size_t pick = 0;  // the pair whose panel partition is closest to 64 SMs
        for (size_t k = 0; k < h->ov.pairs.size(); ++k)
            if (abs(h->ov.pairs[k].nsmP - 64) < abs(h->ov.pairs[pick].nsmP - 64)) pick = k;
        const auto& pr = h->ov.pairs[pick];
        s_panel = pr.sP; s_rest = pr.sP2; s_side = pr.sP3; s_upd = pr.sU;
        nsm_p = pr.nsmP; nsm_u = pr.nsmU;
    }
    cudaStream_t s_comm = g->s_comm;
    MPQR_CUDA(cudaEventRecord(g->ev_start, st));
    MPQR_CUDA(cudaStreamWaitEvent(s_panel, g->ev_start, 0));
    MPQR_CUDA(cudaStreamWaitEvent(s_upd, g->ev_start, 0));
    MPQR_CUDA(cudaStreamWaitEvent(s_comm, g->ev_start, 0));
    std::vector<char> did_next(nblk, 0);
    for (int c0 = 0, b = 0; c0 < h->kmax; c0 += nb, ++b) {
        const int c1 = (c0 + nb < h->kmax) ? c0 + nb : h->kmax;
        const int owner = b % P, Dblk = m - c0;
        void* Y16 = g->YW16[b & 1];
        void* W16 = (char*)g->YW16[b & 1] + (size_t)m * g->ldw * 2;
        BlockCtx c{};
        c.A = dA; c.lda = lda; c.Ah = g->Ah; c.ldh = g->ldh;
        c.Y16 = Y16; c.ldy = g->ldw; c.W16 = W16; c.ldw = g->ldw;
        const int acol0 = (b / P) * nb;                    // local column of global column c0 on the owner
        const int bw_full = (c0 + nb <= n) ? nb : n - c0;  // all columns of global block b (reflectors may stop earlier: m < n)
        if (owner == g->rank) {
            c.acol0 = acol0;
            if (b >= 2) {
                MPQR_CUDA(cudaStreamWaitEvent(s_panel, g->ev_far[b - 2], 0));
                if (did_next[b - 2]) MPQR_CUDA(cudaStreamWaitEvent(s_panel, g->ev_next[b - 2], 0));
            }
            // the block's own trailing columns end where the next local block (a later global block) begins, so a
            // spill past the last in-block column would hit live data unless it is the physical end of the local matrix
            const int end_ok = (acol0 + (c1 - c0) == g->nloc);
            BlockCtx cp = c;
            cp.chain_side = s_side;
            const bool two = s_rest && !h->ov.ev_rest.empty() && !h->ov.ev_acc.empty();
            if (two) {
                // in-block rest updates and the WY accumulation (which rewrites W_p in place, so it must follow the rest
                // update that reads it) on the partition's second stream
                cp.rest_stream = s_rest; cp.rest_S32 = h->S32r; cp.rest_S16 = h->S16r; cp.rest_ev = h->ov.ev_rest.data();
                cp.acc_stream = s_rest; cp.acc_sms = nsm_p; cp.acc_S32 = h->S32r; cp.acc_S16 = h->S16r; cp.acc_ev = h->ov.ev_acc.data();
            }
            {
                SmBudget budget(nsm_p);
                MPQR_TRY(block_phase(h, cp, c0, c1, end_ok, s_panel));
            }
            if (two) {
                MPQR_CUDA(cudaEventRecord(h->ov.ev_accdone, s_rest));
                MPQR_CUDA(cudaStreamWaitEvent(s_panel, h->ov.ev_accdone, 0));
            }
            MPQR_CUDA(cudaEventRecord(g->ev_bp[b], s_panel));
            MPQR_CUDA(cudaStreamWaitEvent(s_comm, g->ev_bp[b], 0));
        }
        if (b >= 2) {  // the receivers overwrite staging[b & 1]: its readers of block b-2 must be done
            MPQR_CUDA(cudaStreamWaitEvent(s_comm, g->ev_far[b - 2], 0));
            if (did_next[b - 2]) MPQR_CUDA(cudaStreamWaitEvent(s_comm, g->ev_next[b - 2], 0));
        }
        // Y and W of block b to everyone (rows c0..m of each staging half are contiguous)
        const size_t bytes = (size_t)Dblk * g->ldw * 2;
        MPQR_NCCL(g->api, g->api->Broadcast(Y16, Y16, bytes, kNcclChar, owner, g->comm, s_comm));
        MPQR_NCCL(g->api, g->api->Broadcast(W16, W16, bytes, kNcclChar, owner, g->comm, s_comm));
        MPQR_CUDA(cudaEventRecord(g->ev_bc[b], s_comm));
        // ---- updates with block b
        // first local column that belongs to a global block > b; on the owner of a block whose reflectors stop before
        // its last column (the last block of a wide matrix whose m is not a multiple of nb), the rest of that block first
        const int li0 = (b >= g->rank) ? (b - g->rank) / P + 1 : 0;
        int afar = li0 * nb;
        if (owner == g->rank && (c1 - c0) < bw_full) afar = acol0 + (c1 - c0);
        BlockCtx cu = c;
        cu.S32 = h->S32u; cu.S16 = h->S16u;
        const bool has_next = (b + 1 < nblk);
        if (has_next && (b + 1) % P == g->rank && afar < g->nloc) {
            // this rank factors block b+1 next: its columns first, on the panel stream (with that stream's GEMM scratch)
            const int lnext0 = ((b + 1) / P) * nb;
            const int wnext = ((b + 2) * nb <= n) ? nb : n - (b + 1) * nb;
            MPQR_CUDA(cudaStreamWaitEvent(s_panel, g->ev_bc[b], 0));
            if (b >= 1) MPQR_CUDA(cudaStreamWaitEvent(s_panel, g->ev_far[b - 1], 0));
            {
                SmBudget budget(nsm_p);
                MPQR_TRY(far_update(h, c, c0, c1, lnext0, wnext, s_panel));
            }
            MPQR_CUDA(cudaEventRecord(g->ev_next[b], s_panel));
            did_next[b] = 1;
            afar = lnext0 + wnext;
        }
        MPQR_CUDA(cudaStreamWaitEvent(s_upd, g->ev_bc[b], 0));
        if (afar < g->nloc) {
            // (these columns were last written by far_rest(b-1) on this stream, or by far_next(b-1) / bp(b) on the panel stream)
            if (b >= 1 && did_next[b - 1]) MPQR_CUDA(cudaStreamWaitEvent(s_upd, g->ev_next[b - 1], 0));
            if (owner == g->rank) MPQR_CUDA(cudaStreamWaitEvent(s_upd, g->ev_bp[b], 0));
            SmBudget budget(nsm_u);
            MPQR_TRY(far_update(h, cu, c0, c1, afar, g->nloc - afar, s_upd));
        }
        MPQR_CUDA(cudaEventRecord(g->ev_far[b], s_upd));
    }
    // join
    MPQR_CUDA(cudaStreamWaitEvent(st, g->ev_far[nblk - 1], 0));
    MPQR_CUDA(cudaStreamWaitEvent(st, g->ev_bc[nblk - 1], 0));
    for (int b = 0; b < nblk; ++b) {
        if (did_next[b]) MPQR_CUDA(cudaStreamWaitEvent(st, g->ev_next[b], 0));
        if (b % P == g->rank) MPQR_CUDA(cudaStreamWaitEvent(st, g->ev_bp[b], 0));
    }
    h->factored = true;
    (void)n;
    return MPQR_OK;
}

// ---------------------------------------------------------------------------------------- multi-GPU TSQR
// 1-D ROW-block layout, rank p owns m_local rows.  Local TSQR -> R_p (n x n) and thin
// Q_p; ncclAllGather of the R_p (n*n*4 bytes per rank, the ONLY exchange); every rank factors the same
// stacked (P*n) x n matrix, rank 0's R is broadcast (the copies agree only to rounding: atomics), and each rank multiplies its
// Q_p by its own n x n slice of the stack's Q.  ts_qr's tree (python/ca_qr.py:36-41) with P leaves.
int mpqr_mg_tsqr_create(mpqr_handle** out, int rank, int nranks, const void* uid) {
    if (!out || !uid || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("mpqr_mg_tsqr_create: bad arguments");
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));  // no CPU fallback: fails without an sm_100 device
    mpqr_handle* h = new mpqr_handle();
    MgState* g = new MgState();
    h->mg = g;
    g->rank = rank;
    g->nranks = nranks;
    int rc = load_nccl(&g->api);
    if (rc == MPQR_OK) {
        NcclUid id;
        memcpy(&id, uid, sizeof(id));
        int nr = g->api->CommInitRank(&g->comm, nranks, id, rank);
        if (nr != 0) {
            set_error("ncclCommInitRank failed: %d", nr);
            rc = MPQR_ENCCL;
        }
    }
    if (rc != MPQR_OK) {
        mpqr_destroy(h);
        return rc;
    }
    *out = h;
    return MPQR_OK;
}

int mpqr_mg_tsqr_device(mpqr_handle* h, const float* dA_local, long lda, long m_local, int n, float* dQ_local, long ldq,
                        float* dR, long ldr, void* stream) {
    if (!h || !h->mg || !dA_local || !dR || n < 1 || m_local < n || lda < n || ldr < n || (dQ_local && ldq < n)) {
        set_error("mpqr_mg_tsqr_device: bad arguments (every rank needs m_local >= n)");
        return MPQR_EINVAL;
    }
    MgState* g = (MgState*)h->mg;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = g->nranks;
    if (P == 1) return mpqr_tsqr_device(dA_local, lda, m_local, n, dQ_local, ldq, dR, ldr, stream);
    float *rloc = nullptr, *rstack = nullptr, *qstack = nullptr, *qtmp = nullptr;
    const size_t nn = (size_t)n * n;
    int rc = MPQR_OK;
    do {
        if ((rc = g->tq_get(0, nn * sizeof(float), &rloc)) || (rc = g->tq_get(1, nn * P * sizeof(float), &rstack))) break;
        if (dQ_local && ((rc = g->tq_get(2, nn * P * sizeof(float), &qstack)) ||
                         (rc = g->tq_get(3, (size_t)m_local * n * sizeof(float), &qtmp)))) break;
        // local leaf: R_p and (optionally) the thin Q_p, kept aside until the tree is known
        if ((rc = mpqr_tsqr_device(dA_local, lda, m_local, n, dQ_local ? qtmp : nullptr, n, rloc, n, st))) break;
        int nr = g->api->AllGather(rloc, rstack, nn * sizeof(float), kNcclChar, g->comm, st);
        if (nr != 0) { set_error("ncclAllGather failed: %d", nr); rc = MPQR_ENCCL; break; }
        // every rank factors the same (P n) x n stack
        if ((rc = mpqr_tsqr_device(rstack, n, (long)P * n, n, dQ_local ? qstack : nullptr, n, dR, ldr, st))) break;
        // The stack is factored redundantly, but the in-panel reductions use atomics, so the copies agree only to
        // rounding: rank 0's R is made THE R on every rank (256 KB); each rank's Q rows stay consistent with its
        // own copy to the same rounding level.
        const size_t rowb = (size_t)n * sizeof(float);
        if (cudaMemcpy2DAsync(rloc, rowb, dR, (size_t)ldr * sizeof(float), rowb, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("mpqr_mg_tsqr_device: copy failed"); rc = MPQR_ECUDA; break;
        }
        nr = g->api->Broadcast(rloc, rloc, nn * sizeof(float), kNcclChar, 0, g->comm, st);
        if (nr != 0) { set_error("ncclBroadcast failed: %d", nr); rc = MPQR_ENCCL; break; }
        if (cudaMemcpy2DAsync(dR, (size_t)ldr * sizeof(float), rloc, rowb, rowb, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("mpqr_mg_tsqr_device: copy failed"); rc = MPQR_ECUDA; break;
        }
        if (dQ_local) {
            rc = sgemm_nn_store(qtmp, n, qstack + (size_t)g->rank * nn, n, dQ_local, ldq, (int)m_local, n, n, st);
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == MPQR_OK && e != cudaSuccess) { set_error("mpqr_mg_tsqr_device: %s", cudaGetErrorString(e)); rc = MPQR_ECUDA; }
    return rc;
}

}  // extern "C"

namespace mpqr {
void mg_destroy(void* state) {
    MgState* g = (MgState*)state;
    if (!g) return;
    if (g->comm && g->api) g->api->CommDestroy(g->comm);
    for (auto e : g->ev_bp) cudaEventDestroy(e);
    for (auto e : g->ev_bc) cudaEventDestroy(e);
    for (auto e : g->ev_far) cudaEventDestroy(e);
    for (auto e : g->ev_next) cudaEventDestroy(e);
    if (g->ev_start) cudaEventDestroy(g->ev_start);
    if (g->s_comm) cudaStreamDestroy(g->s_comm);
    if (g->own_streams) { if (g->s_panel) cudaStreamDestroy(g->s_panel); if (g->s_upd) cudaStreamDestroy(g->s_upd); }
    for (float* p : g->tq) cudaFree(p);
    delete g;
}
}  // namespace mpqr
