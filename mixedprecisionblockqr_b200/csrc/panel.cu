// panel.cu — Householder panel factorisation + compact-WY (sm_100a), three-level scheme.
//
// Replaces the reference's HOST panel factorisation h_householder_qr (Cuda/qr.cu:198-293, one
// CPU thread, full-matrix PCIe round trip per panel, :1080-1082/:1215) and the 3r+2 launches of
// dev_wy_transform (Cuda/qr.cu:535-600, K1-K4 of SURVEY 2.4).
//
//   level 0  panel_block_kernel<B,RPT>: B (16 or 32) columns x D rows, REGISTER resident.  The
//            D x B block is spread by rows over ONE thread-block cluster (<= 16 CTAs x 512
//            threads x RPT rows); every thread keeps its rows' B values in registers for all B
//            reflector steps, so HBM sees one read and one write of the block.  Per column ONE
//            reduction: the dots g_j = u^T a_j give the norm (j = k), the update coefficients
//            v^T a_j = g_j + s*mu*a_kj (j > k) and the Gram entries y_j^T y_k (j < k) at once;
//            the rank-1 update of step k-1 and the dots of step k share one pass.  Reduction:
//            warp transpose-reduce (shuffles) -> shared memory -> all-gather of the CTA sums
//            through distributed shared memory with st.async + mbarrier complete_tx (no
//            cluster-wide barrier, no L2 round trip inside the column loop).  Tail: T of the
//            block by the larft recurrence, W = Y T, packed output, FP32 + FP16/BF16 copies.
//   level 1  inside an r-wide panel (r <= 128) the blocks are combined right-looking in FP32:
//            S = W_j^T A_rest (inpanel_s_kernel), A_rest -= Y_j S (inpanel_u_kernel).
//   level 2  T of the whole panel from the Gram matrix, T = (striu(Y^T Y) + I/2)^-1
//            (tinv_kernel, recursive doubling in one CTA), then W = Y T.
//            Mixed-precision driver: Gram and W on tcgen05 from the SAME 16-bit Y the trailing
//            GEMMs use (so I - Y16 T Y16^T is orthogonal to FP32-accumulate accuracy);
//            FP32 driver: SIMT GEMMs.
//
// Conventions mirrored from the reference: sign = (u0 >= 0) ? +1 : -1
// (:229-235); zero column => reflector skipped (:242-244); unit vector w (beta = 2) stored
// one row below the diagonal (:283-285); R_kk = -sign*||u||.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpqr {

int launch_panel_legacy(const PanelArgs& a, cudaStream_t stream, long* launches);
int sgemm_nn_store(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K,
                   cudaStream_t stream);

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int CSMAX = 16;
constexpr int SLD = 256;   // leading dimension of the S replicas (in-panel rest + the next panel's columns)
constexpr int NREP = 8;    // S replicas (spreads the atomics over L2 slices)
constexpr int RMAX = kPanelMaxWidth;

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// remote 16-byte store that signals 16 bytes on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// The awaited data are st.async writes into THIS CTA's shared memory, completed through the barrier's
// tx-count: the default (CTA-scope acquire) wait is sufficient and avoids the L1 invalidate (CCTL.IVALL)
// that a cluster-scope acquire emits on every step.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
        ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void store16(void* base, long idx, float v, int bf16) {
    if (bf16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int bf16) {
    if (bf16) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    }
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// One stage of the warp transpose-reduce: N values per lane -> N/2, partner = lane ^ OFF.
template <int N, int OFF>
__device__ __forceinline__ void tr_stage(float* acc, int lane) {
    const bool hi = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = hi ? acc[i] : acc[i + N / 2];
        const float keep = hi ? acc[i + N / 2] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}
// After the call lane l holds in acc[0] the warp sum of column (B == 32 ? l : l >> 1).
template <int B>
__device__ __forceinline__ void warp_transpose_reduce(float* acc, int lane) {
    if (B == 32) {
        tr_stage<32, 16>(acc, lane);
        tr_stage<16, 8>(acc, lane);
        tr_stage<8, 4>(acc, lane);
        tr_stage<4, 2>(acc, lane);
        tr_stage<2, 1>(acc, lane);
    } else {
        tr_stage<16, 16>(acc, lane);
        tr_stage<8, 8>(acc, lane);
        tr_stage<4, 4>(acc, lane);
        tr_stage<2, 2>(acc, lane);
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
    }
}

// T = (striu(G) + I/2)^-1 in shared memory, all NT threads: 8 x 8 diagonal blocks by back
// substitution from registers (x_c = 2, x_t = -2 sum_{u=t+1..c} G[t][u] x_u), then pairs of
// blocks are merged, T12 = -T11 (G12 T22), 8 -> 16 -> 32.  gt: B x LD, strictly upper = G.
template <int B, int LD>
__device__ __forceinline__ void tinv_smem(float (*gt)[LD], float* xs, int tid) {
    float xv[8];
    const int b0 = tid & ~7, cc = tid & 7;
    if (tid < B) {
        float g[8][8];
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int u2 = t + 1; u2 < 8; ++u2) g[t][u2] = gt[b0 + t][b0 + u2];
#pragma unroll
        for (int t = 7; t >= 0; --t) {
            float sacc = 0.f;
#pragma unroll
            for (int u2 = t + 1; u2 < 8; ++u2)
                if (u2 <= cc) sacc = fmaf(g[t][u2], xv[u2], sacc);
            xv[t] = (t == cc) ? 2.f : ((t < cc) ? -2.f * sacc : 0.f);
        }
    }
    __syncthreads();  // every column has read G before anyone overwrites it
    if (tid < B) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (t <= cc) gt[b0 + t][b0 + cc] = xv[t];
    }
    __syncthreads();
#pragma unroll
    for (int h = 8; h < B; h *= 2) {
        // pairs of blocks [o, o+h) and [o+h, o+2h); B*h/2 <= NT outputs per matrix product
        const int pr = tid / (h * h), e = tid - pr * h * h;
        const int i = e / h, j = e - i * h, o = pr * 2 * h;
        const bool on = tid < (B / (2 * h)) * h * h;
        if (on) {
            float s0 = 0.f, s1 = 0.f;  // zeros below the diagonals: fixed trip counts
#pragma unroll
            for (int u2 = 0; u2 < h; u2 += 2) {
                s0 = fmaf(gt[o + i][o + h + u2], gt[o + h + u2][o + h + j], s0);
                s1 = fmaf(gt[o + i][o + h + u2 + 1], gt[o + h + u2 + 1][o + h + j], s1);
            }
            xs[tid] = s0 + s1;  // X = G12 T22, element (pr, i, j)
        }
        __syncthreads();
        if (on) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int u2 = 0; u2 < h; u2 += 2) {
                s0 = fmaf(gt[o + i][o + u2], xs[pr * h * h + u2 * h + j], s0);
                s1 = fmaf(gt[o + i][o + u2 + 1], xs[pr * h * h + (u2 + 1) * h + j], s1);
            }
            gt[o + i][o + h + j] = -(s0 + s1);
        }
        __syncthreads();
    }
}

struct Out32 {
    float* p;   // element (block row 0, block column 0); null = not wanted
    long ld;
    int zrows;  // rows above the block that must be zero-filled (p - zrows*ld is their first row)
};
struct Out16 {
    void* p;
    long ld;
    int zrows;
};

struct BlockArgs {
    float* A;   // element (block row 0, block column 0) of the packed FP32 master
    long lda;
    int D;      // rows of the block (m - first row)
    int bw;     // columns (<= B)
    Out32 Y32, W32;
    Out16 Y16, W16;
    int bf16;
    float* T;   // bw x bw (ldt) upper triangular block T, or null
    int ldt;
    float* zero_buf;  // optional buffer to clear (S replicas of the following in-panel update)
    int zero_n;
    long long* dbg;   // optional phase timers (tools/panel_probe.py)
    int defer_out;    // 1: the packed factor below the first 32 rows and the 16-bit Y are written later, per panel, by
                      // panel_finalize_kernel from the FP32 Y (the kernel then only stores Y32 and its top 32 rows of A)
};

template <int B>
__device__ __forceinline__ void load_row(const float* p, float (&x)[B], int bw, bool vec) {
    if (vec && bw == B) {
#pragma unroll
        for (int q = 0; q < B / 4; ++q) {
            float4 v = __ldcg(reinterpret_cast<const float4*>(p) + q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < B; ++c) x[c] = (c < bw) ? __ldcg(p + c) : 0.f;
    }
}
template <int B>
__device__ __forceinline__ void store_row32(float* p, const float (&x)[B], int bw, bool vec) {
    if (vec && bw == B) {
#pragma unroll
        for (int q = 0; q < B / 4; ++q)
            reinterpret_cast<float4*>(p)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < B; ++c)
            if (c < bw) p[c] = x[c];
    }
}
template <int B>
__device__ __forceinline__ void store_row16(void* p, const float (&x)[B], int bw, bool vec, int bf16) {
    if (vec && bw == B) {
#pragma unroll
        for (int q = 0; q < B / 8; ++q)
            reinterpret_cast<uint4*>(p)[q] = make_uint4(pack16(x[8 * q], x[8 * q + 1], bf16), pack16(x[8 * q + 2], x[8 * q + 3], bf16),
                                                        pack16(x[8 * q + 4], x[8 * q + 5], bf16), pack16(x[8 * q + 6], x[8 * q + 7], bf16));
    } else {
#pragma unroll
        for (int c = 0; c < B; ++c)
            if (c < bw) store16(p, c, x[c], bf16);
    }
}

// ---- warp-cooperative, coalesced transfers between global rows and "one row per lane" registers.
// A lane that moves its own row touches 32 different cache lines per instruction (a 64-byte row
// segment per lane): ncu showed 10 us of load + 10 us of stores per 16-column block at D = 32768,
// i.e. 32 GB/s per SM.  Here CH lanes share a row (CH = 16-byte chunks per row), so one
// instruction covers 32/CH whole rows; the transposition to/from the row-per-lane layout goes through
// a warp-private shared tile whose chunks are XOR-swizzled (conflict-free both ways).
//   tile: 32 * CH uint4 per warp.   g: element (first row of the group, first column), pitch in BYTES.
template <int CH>
__device__ __forceinline__ int tile_swz(int row) { return (row / (8 / CH)) % CH; }

template <int CH>
__device__ __forceinline__ void warp_tile_in(const char* g, size_t pitch, int nvalid, uint4 (&v)[CH], uint4* tile, int lane) {
    constexpr int RPI = 32 / CH;  // rows per instruction
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int r = k * RPI + lane / CH, q = lane % CH;
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (r < nvalid) t = __ldcg(reinterpret_cast<const uint4*>(g + (size_t)r * pitch) + q);
        tile[r * CH + (q ^ tile_swz<CH>(r))] = t;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < CH; ++q) v[q] = tile[lane * CH + (q ^ tile_swz<CH>(lane))];
    __syncwarp();
}
// The same in two halves, so that a caller can put ALL its rows' global loads in flight before the first
// transposition (the __syncwarp()s of one fused call would serialise one L2 round trip per 32-row group).
template <int CH>
__device__ __forceinline__ void warp_tile_fetch(const char* g, size_t pitch, int nvalid, uint4 (&t)[CH], int lane) {
    constexpr int RPI = 32 / CH;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int r = k * RPI + lane / CH, q = lane % CH;
        t[k] = make_uint4(0u, 0u, 0u, 0u);
        if (r < nvalid) t[k] = __ldcg(reinterpret_cast<const uint4*>(g + (size_t)r * pitch) + q);
    }
}
template <int CH>
__device__ __forceinline__ void warp_tile_transpose(const uint4 (&t)[CH], uint4 (&v)[CH], uint4* tile, int lane) {
    constexpr int RPI = 32 / CH;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int r = k * RPI + lane / CH, q = lane % CH;
        tile[r * CH + (q ^ tile_swz<CH>(r))] = t[k];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < CH; ++q) v[q] = tile[lane * CH + (q ^ tile_swz<CH>(lane))];
    __syncwarp();
}

template <int CH>
__device__ __forceinline__ void warp_tile_out(char* g, size_t pitch, int nvalid, const uint4 (&v)[CH], uint4* tile, int lane) {
    constexpr int RPI = 32 / CH;
#pragma unroll
    for (int q = 0; q < CH; ++q) tile[lane * CH + (q ^ tile_swz<CH>(lane))] = v[q];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int r = k * RPI + lane / CH, q = lane % CH;
        if (r < nvalid) reinterpret_cast<uint4*>(g + (size_t)r * pitch)[q] = tile[r * CH + (q ^ tile_swz<CH>(r))];
    }
    __syncwarp();
}
template <int B>
__device__ __forceinline__ void rows_in_f32(const float* g, long ld, int nvalid, float (&x)[B], uint4* tile, int lane) {
    uint4 v[B / 4];
    warp_tile_in<B / 4>(reinterpret_cast<const char*>(g), (size_t)ld * 4, nvalid, v, tile, lane);
#pragma unroll
    for (int q = 0; q < B / 4; ++q) {
        x[4 * q] = __uint_as_float(v[q].x); x[4 * q + 1] = __uint_as_float(v[q].y);
        x[4 * q + 2] = __uint_as_float(v[q].z); x[4 * q + 3] = __uint_as_float(v[q].w);
    }
}
template <int B>
__device__ __forceinline__ void rows_out_f32(float* g, long ld, int nvalid, const float (&x)[B], uint4* tile, int lane) {
    uint4 v[B / 4];
#pragma unroll
    for (int q = 0; q < B / 4; ++q)
        v[q] = make_uint4(__float_as_uint(x[4 * q]), __float_as_uint(x[4 * q + 1]), __float_as_uint(x[4 * q + 2]), __float_as_uint(x[4 * q + 3]));
    warp_tile_out<B / 4>(reinterpret_cast<char*>(g), (size_t)ld * 4, nvalid, v, tile, lane);
}
template <int B>
__device__ __forceinline__ void rows_out_16(void* g, long ld, int nvalid, const float (&x)[B], uint4* tile, int lane, int bf16) {
    uint4 v[B / 8];
#pragma unroll
    for (int q = 0; q < B / 8; ++q)
        v[q] = make_uint4(pack16(x[8 * q], x[8 * q + 1], bf16), pack16(x[8 * q + 2], x[8 * q + 3], bf16),
                          pack16(x[8 * q + 4], x[8 * q + 5], bf16), pack16(x[8 * q + 6], x[8 * q + 7], bf16));
    warp_tile_out<B / 8>(reinterpret_cast<char*>(g), (size_t)ld * 2, nvalid, v, tile, lane);
}

// dbg slots: 0 pass, 1 shuffle tree, 2 smem + CTA barrier, 3 CTA sum + send, 4 exchange wait, 5 gather+scalars,
// 6 load, 7 tail, 8 steps, 9 CS, 10 RPT, 11 B
#define PROF_MARK(slot)                     \
    if (prof) {                             \
        long long t__ = clock64();          \
        pacc[slot] += t__ - tprev;          \
        tprev = t__;                        \
    }

// The reflector steps of one B-column register block (rows rbase + u*NT of the cluster's slab; the
// block's diagonal sits at row `roff` of the slab: 0, or 16 for the second half of a double block).
template <int B>
struct StepMem {
    float (*red)[NW][B];
    float (*prow)[B];
    float (*slot)[CSMAX][B];
    float (*pslot)[B];
    float (*tauS)[B];
    float (*gt)[B + 4];
    float* diag;
    uint64_t* mbar;
};
struct StepCtx {
    int tid, lane, warp, CS, rbase, bw, kr, roff;
    unsigned crank;
    bool prof;
    long long* pacc;
    long long* tprev_p;
};

template <int B, int RPT>
__device__ __forceinline__ void factor_steps(float (&x)[RPT][B], const StepMem<B>& M, const StepCtx& c) {
    constexpr int LB = (B == 32) ? 0 : 1;  // lane -> column shift after the transpose-reduce
    float (*red)[NW][B] = M.red;
    float (*prow)[B] = M.prow;
    float (*slot)[CSMAX][B] = M.slot;
    float (*pslot)[B] = M.pslot;
    float (*tauS)[B] = M.tauS;
    float (*gt)[B + 4] = M.gt;
    float* diag = M.diag;
    uint64_t* mbar = M.mbar;
    const int tid = c.tid, lane = c.lane, warp = c.warp, CS = c.CS, rbase = c.rbase, bw = c.bw, kr = c.kr, roff = c.roff;
    const unsigned crank = c.crank;
    const bool prof = c.prof;
    long long* pacc = c.pacc;
    long long& tprev = *c.tprev_p;
    const uint32_t tx_bytes = (uint32_t)(CS + 1) * B * 4;
    const int mypos = (lane >> LB) & (B - 1);  // position this lane owns after the transpose-reduce

    // The step loop is ROLLED (a fully unrolled body is ~300 KB of straight-line SASS and the
    // kernel then starves on instruction fetch: ncu showed 70 % stall_no_inst).  Registers
    // cannot be indexed dynamically, so the block is ROTATED instead: the current column always
    // sits at position 0, and the rank-1 update writes position p into p-1
    // (x[p-1] = x[p] - v*tau[p], one FFMA does update + rotation); the finished column re-enters
    // at position B-1.  After B steps every column is back at its own position.
    // Position p at step s holds column (s + p) mod B: live for p < B - s, finished otherwise.
#pragma unroll 1
    for (int s = 0; s < B; ++s) {
        const int par = s & 1;
        const bool active = s < kr;  // ragged blocks: the remaining steps only rotate
        // ---- dots of the current column with every position (norm, update coefficients, Gram)
        float acc[B];
        float xs[RPT];
        if (crank == 0 && tid == s + roff) {
#pragma unroll
            for (int q = 0; q < B / 4; ++q)  // pivot row (row s is u = 0 of thread s), positions
                *reinterpret_cast<float4*>(&prow[par][4 * q]) = make_float4(x[0][4 * q], x[0][4 * q + 1], x[0][4 * q + 2], x[0][4 * q + 3]);
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int i = rbase + u * NT;
            xs[u] = (i >= s + roff) ? x[u][0] : 0.f;  // rows above the diagonal hold R entries
            if (u == 0) {
#pragma unroll
                for (int p2 = 0; p2 < B; ++p2) acc[p2] = xs[0] * x[0][p2];
            } else {
#pragma unroll
                for (int p2 = 0; p2 < B; ++p2) acc[p2] = fmaf(xs[u], x[u][p2], acc[p2]);
            }
        }
        PROF_MARK(0);
        warp_transpose_reduce<B>(acc, lane);
        PROF_MARK(1);
        if (LB == 0 || (lane & 1) == 0) red[par][warp][mypos] = acc[0];
        if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&mbar[par], tx_bytes);
        __syncthreads();
        PROF_MARK(2);
        float g, pv;
        if (CS > 1) {
            {
                // one lane per (peer, 4-position chunk): the lane sums its chunk over the warps and ships
                // 16 bytes to the peer's slot (+ the pivot row chunk from CTA 0)
                constexpr int CH = B / 4;
                const int o = warp * 32 + lane;
                if (o < CS * CH) {
                    const unsigned peer = (unsigned)(o / CH);
                    const int ch = o % CH;
                    float4 cs4 = *reinterpret_cast<const float4*>(&red[par][0][4 * ch]);
#pragma unroll
                    for (int w = 1; w < NW; ++w) {
                        const float4 r4 = *reinterpret_cast<const float4*>(&red[par][w][4 * ch]);
                        cs4.x += r4.x; cs4.y += r4.y; cs4.z += r4.z; cs4.w += r4.w;
                    }
                    const uint32_t rbar = map_to_cta(smem_addr(&mbar[par]), peer);
                    st_async_v4(map_to_cta(smem_addr(&slot[par][crank][4 * ch]), peer), cs4, rbar);
                    if (crank == 0)
                        st_async_v4(map_to_cta(smem_addr(&pslot[par][4 * ch]), peer),
                                    *reinterpret_cast<const float4*>(&prow[par][4 * ch]), rbar);
                }
            }
            PROF_MARK(3);
            mbar_wait_cluster(&mbar[par], (uint32_t)((s >> 1) & 1));
            PROF_MARK(4);
            // (fully unrolled over the largest cluster with a uniform predicate: the 16 LDS issue back to back instead
            // of one latency-exposed pair per trip of a rolled loop; CS is even)
            float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
            for (int c = 0; c < CSMAX; c += 4) {
                if (c < CS) {
                    g0 += slot[par][c][lane & (B - 1)];
                    g1 += slot[par][c + 1][lane & (B - 1)];
                }
                if (c + 2 < CS) {
                    g2 += slot[par][c + 2][lane & (B - 1)];
                    g3 += slot[par][c + 3][lane & (B - 1)];
                }
            }
            g = (g0 + g1) + (g2 + g3);
            pv = pslot[par][lane & (B - 1)];
        } else {
            float csum = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) csum += red[par][w][lane & (B - 1)];
            PROF_MARK(3);
            g = csum;
            pv = prow[par][lane & (B - 1)];
        }
        // ---- reflector scalars (every lane redundantly): MUFU.RSQ + one Newton step each
        const float gk = __shfl_sync(0xffffffffu, g, 0);
        const float ak = __shfl_sync(0xffffffffu, pv, 0);
        const bool skip = !(gk > 0.f) || !active;
        const float rs = rsqrtf(skip ? 1.f : gk);
        float mu = gk * rs;
        mu = fmaf(0.5f * rs, fmaf(-mu, mu, gk), mu);  // sqrt(gk)
        if (skip) mu = 0.f;
        float smu = (ak >= 0.f) ? mu : -mu;
        const float vn2 = 2.f * mu * (mu + fabsf(ak));
        float rv = rsqrtf(skip ? 1.f : vn2);
        rv = rv * fmaf(-0.5f * vn2, rv * rv, 1.5f);  // 1/sqrt(vn2)
        const float vinv = skip ? 0.f : rv;
        const float inv2 = skip ? 0.f : 2.f * rv * rv;  // 2/vn2
        const int pos = lane & (B - 1);
        const int col = (s + pos) & (B - 1);  // column held at this position
        const float t = fmaf(smu, pv, g);
        if (lane < B) tauS[warp][pos] = (pos >= 1 && pos < B - s && col < bw) ? t * inv2 : 0.f;
        if (warp == 0 && lane < B) {
            if (pos >= B - s) gt[col][s] = t * vinv;        // Gram entry y_col^T y_s (col < s)
            if (pos == 0) diag[s] = skip ? ak : -smu;       // R_ss
        }
        __syncwarp();
        // ---- rank-1 update fused with the rotation
        float tau[B];
#pragma unroll
        for (int q = 0; q < B / 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(&tauS[warp][4 * q]);
            tau[4 * q] = t4.x; tau[4 * q + 1] = t4.y; tau[4 * q + 2] = t4.z; tau[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int i = rbase + u * NT;
            const float vi = (i == s + roff) ? xs[u] + smu : xs[u];
            const float fin = (i >= s + roff && active) ? vi * vinv : x[u][0];  // w in place (unshifted); R entries above
#pragma unroll
            for (int p2 = 1; p2 < B; ++p2) x[u][p2 - 1] = fmaf(-vi, tau[p2], x[u][p2]);
            x[u][B - 1] = fin;
        }
        PROF_MARK(5);
    }
    }

template <int B, int RPT>
__global__ void __launch_bounds__(NT, 1) panel_block_kernel(BlockArgs a, int CS) {
    __shared__ __align__(16) float red[2][NW][B];
    __shared__ __align__(16) float prow[2][B];
    __shared__ __align__(16) float slot[2][CSMAX][B];
    __shared__ __align__(16) float pslot[2][B];
    __shared__ __align__(16) float tauS[NW][B];
    __shared__ __align__(16) float gt[B][B + 4];
    __shared__ float diag[B];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ __align__(16) uint4 tiles[NW][32 * (B / 4)];  // warp-private transposition tiles (coalesced I/O)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = (CS > 1) ? cluster_ctarank() : 0u;
    const int D = a.D, bw = a.bw;
    const int kr = bw < D ? bw : D;  // reflectors of this block
    const int rbase = (int)crank * (NT * RPT) + tid;  // row of u = 0; row(u) = rbase + u*NT
    const long lda = a.lda;
    const bool vecA = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);

    const bool prof = (a.dbg != nullptr) && blockIdx.x == 0 && tid == 0;
    long long pacc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = prof ? clock64() : 0;

    pdl_launch_dependents();
    for (int idx = tid; idx < B * (B + 4); idx += NT) (&gt[0][0])[idx] = 0.f;
    if (CS > 1) {
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();  // peers must not signal a barrier that is not initialised yet
    }
    pdl_wait();  // everything above overlapped the predecessor's tail; its results are visible from here on

    float x[RPT][B];
    const int wrow0 = rbase - lane;  // first row of this warp's 32-row group for u = 0
    if (vecA && bw == B) {
        // all RPT groups' loads in flight first (RPT * B / 4 independent 16-byte loads per lane), then the transpositions
        uint4 t[RPT][B / 4];
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int g0 = wrow0 + u * NT;
            warp_tile_fetch<B / 4>(reinterpret_cast<const char*>(a.A + (size_t)g0 * lda), (size_t)lda * 4, D - g0, t[u], lane);  // rows >= D read as zero
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            uint4 v[B / 4];
            warp_tile_transpose<B / 4>(t[u], v, tiles[warp], lane);
#pragma unroll
            for (int q = 0; q < B / 4; ++q) {
                x[u][4 * q] = __uint_as_float(v[q].x); x[u][4 * q + 1] = __uint_as_float(v[q].y);
                x[u][4 * q + 2] = __uint_as_float(v[q].z); x[u][4 * q + 3] = __uint_as_float(v[q].w);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int i = rbase + u * NT;
            if (i < D) {
                load_row<B>(a.A + (size_t)i * lda, x[u], bw, vecA);
            } else {
#pragma unroll
                for (int c = 0; c < B; ++c) x[u][c] = 0.f;
            }
        }
    }
    if (a.zero_buf) {
        const int nthr = CS * NT;
        for (int idx = (int)crank * NT + tid; idx < a.zero_n; idx += nthr) a.zero_buf[idx] = 0.f;
    }
    __syncthreads();
    PROF_MARK(6);

    {
        StepMem<B> M{red, prow, slot, pslot, tauS, gt, diag, mbar};
        StepCtx sc{tid, lane, warp, CS, rbase, bw, kr, 0, crank, prof, pacc, &tprev};
        factor_steps<B, RPT>(x, M, sc);
    }
    __syncthreads();

    // ---- T of the block, T = (striu(G) + I/2)^-1 (recursive doubling in shared memory)
    tinv_smem<B, B + 4>(gt, &red[0][0][0], tid);

    // ---- outputs
    const bool want_w = a.W32.p || a.W16.p;
    const bool vecY32 = a.Y32.p && ((a.Y32.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y32.p) & 15) == 0);
    const bool vecW32 = a.W32.p && ((a.W32.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.W32.p) & 15) == 0);
    const bool vecY16 = a.Y16.p && ((a.Y16.ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.Y16.p) & 15) == 0);
    const bool vecW16 = a.W16.p && ((a.W16.ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.W16.p) & 15) == 0);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        const int i = rbase + u * NT;
        const int g0 = wrow0 + u * NT;          // first row of the warp's group
        const bool full = (bw == B) && g0 >= B;  // the whole group lies below the block's triangle
        // packed factor: R above the diagonal, R_kk on it, w shifted one row down
        if (full && vecA) {
            if (!a.defer_out) rows_out_f32<B>(a.A + (size_t)(g0 + 1) * lda, lda, D - g0, x[u], tiles[warp], lane);
        } else if (i < D) {
            if (i >= bw) {
                store_row32<B>(a.A + (size_t)(i + 1) * lda, x[u], bw, vecA);
            } else {
#pragma unroll
                for (int c = 0; c < B; ++c)
                    if (c < bw) a.A[(size_t)(i + (i >= c ? 1 : 0)) * lda + c] = x[u][c];
                a.A[(size_t)i * lda + i] = diag[i];  // i < bw: R_ii (the loop above put w_ii one row below)
            }
        }
        // Y: zero strictly above the diagonal and for columns without reflector
        if (i < B) {
#pragma unroll
            for (int c = 0; c < B; ++c)
                if (i < c) x[u][c] = 0.f;
        }
        if (a.Y32.p) {
            if (bw == B && vecY32) rows_out_f32<B>(a.Y32.p + (size_t)g0 * a.Y32.ld, a.Y32.ld, D - g0, x[u], tiles[warp], lane);
            else if (i < D) store_row32<B>(a.Y32.p + (size_t)i * a.Y32.ld, x[u], bw, vecY32);
        }
        if (a.Y16.p && !a.defer_out) {
            if (bw == B && vecY16) rows_out_16<B>((char*)a.Y16.p + (size_t)g0 * a.Y16.ld * 2, a.Y16.ld, D - g0, x[u], tiles[warp], lane, a.bf16);
            else if (i < D) store_row16<B>((char*)a.Y16.p + (size_t)i * a.Y16.ld * 2, x[u], bw, vecY16, a.bf16);
        }
    }
    if (want_w) {
        constexpr int WG = (RPT >= 2) ? 2 : 1;  // rows that share one sweep over T
#pragma unroll
        for (int u0 = 0; u0 < RPT; u0 += WG) {
            float w[WG][B];
#pragma unroll
            for (int g2 = 0; g2 < WG; ++g2)
#pragma unroll
                for (int c = 0; c < B; ++c) w[g2][c] = 0.f;
#pragma unroll
            for (int t = 0; t < B; ++t) {
#pragma unroll
                for (int q = t / 4; q < B / 4; ++q) {
                    const float4 t4 = *reinterpret_cast<const float4*>(&gt[t][4 * q]);  // zero below the diagonal
#pragma unroll
                    for (int g2 = 0; g2 < WG; ++g2) {
                        const float y = x[u0 + g2][t];
                        w[g2][4 * q] = fmaf(y, t4.x, w[g2][4 * q]);
                        w[g2][4 * q + 1] = fmaf(y, t4.y, w[g2][4 * q + 1]);
                        w[g2][4 * q + 2] = fmaf(y, t4.z, w[g2][4 * q + 2]);
                        w[g2][4 * q + 3] = fmaf(y, t4.w, w[g2][4 * q + 3]);
                    }
                }
            }
#pragma unroll
            for (int g2 = 0; g2 < WG; ++g2) {
                const int i = rbase + (u0 + g2) * NT, g0 = wrow0 + (u0 + g2) * NT;
                if (a.W32.p) {
                    if (bw == B && vecW32) rows_out_f32<B>(a.W32.p + (size_t)g0 * a.W32.ld, a.W32.ld, D - g0, w[g2], tiles[warp], lane);
                    else if (i < D) store_row32<B>(a.W32.p + (size_t)i * a.W32.ld, w[g2], bw, vecW32);
                }
                if (a.W16.p) {
                    if (bw == B && vecW16) rows_out_16<B>((char*)a.W16.p + (size_t)g0 * a.W16.ld * 2, a.W16.ld, D - g0, w[g2], tiles[warp], lane, a.bf16);
                    else if (i < D) store_row16<B>((char*)a.W16.p + (size_t)i * a.W16.ld * 2, w[g2], bw, vecW16, a.bf16);
                }
            }
        }
    }
    // rows above the block are structurally zero in the compact outputs
    {
        const int gtid = (int)crank * NT + tid, nthr = CS * NT;
        if (a.Y32.p)
            for (long idx = gtid; idx < (long)a.Y32.zrows * bw; idx += nthr) {
                long rr = idx / bw; int c = (int)(idx - rr * bw);
                a.Y32.p[(rr - a.Y32.zrows) * a.Y32.ld + c] = 0.f;
            }
        if (a.W32.p)
            for (long idx = gtid; idx < (long)a.W32.zrows * bw; idx += nthr) {
                long rr = idx / bw; int c = (int)(idx - rr * bw);
                a.W32.p[(rr - a.W32.zrows) * a.W32.ld + c] = 0.f;
            }
        if (a.Y16.p && !a.defer_out)
            for (long idx = gtid; idx < (long)a.Y16.zrows * bw; idx += nthr) {
                long rr = idx / bw; int c = (int)(idx - rr * bw);
                store16(a.Y16.p, (rr - a.Y16.zrows) * a.Y16.ld + c, 0.f, a.bf16);
            }
        if (a.W16.p)
            for (long idx = gtid; idx < (long)a.W16.zrows * bw; idx += nthr) {
                long rr = idx / bw; int c = (int)(idx - rr * bw);
                store16(a.W16.p, (rr - a.W16.zrows) * a.W16.ld + c, 0.f, a.bf16);
            }
    }
    if (a.T && crank == 0) {
        for (int idx = tid; idx < bw * bw; idx += NT) {
            int t = idx / bw, c = idx - t * bw;
            a.T[(size_t)t * a.ldt + c] = (t <= c && c < kr) ? gt[t][c] : 0.f;
        }
    }
    PROF_MARK(7);
    if (prof) {
        pacc[8] = kr;
        for (int i = 0; i < 9; ++i) a.dbg[i] += pacc[i];
        a.dbg[9] = CS; a.dbg[10] = RPT; a.dbg[11] = B;
    }
    // shared memory must stay alive until no peer can signal into it any more
    if (CS > 1) cluster_sync_all();
}

// ------------------------------------------------------------------ persistent panel chain
// ONE cluster launch factors every 16-column register block of an r-wide panel (panel_chain_kernel).  The rows keep
// their threads for the whole panel (slab row i = crank*NT*RPT + u*NT + tid; block jb's diagonal sits at slab row
// 16*jb, rows above it are inactive), so per block the chain pays neither a launch nor a drain:
//   wait far(jb-2)   the side stream's update of block jb's columns by blocks <= jb-2 (device flag, see below)
//   load X_jb        D x 16 from A (L2 hits: written by the update kernels just before), coalesced through warp tiles
//   near update      X -= Y_{jb-1} T_{jb-1}^T (Y_{jb-1}^T X): Y_{jb-1} comes back from THIS thread's shared-memory spill
//                    (thread-private rows, conflict-free layout), P through one DSMEM all-gather
//   16 reflector steps (factor_steps), T_jb (tinv_smem)
//   outputs          FP32 Y (compact, coalesced), the block's first 32 packed rows of A, T_jb; Y_jb -> shared memory
//   signal           flag_done = base + jb + 1 once every CTA's stores are fenced
// The update of the rest of the panel (blocks jb+2 ..) stays on the device-wide S/U kernels, issued by the host on a
// side stream BEHIND a stream wait on flag_done (cuStreamWaitValue32, or a one-thread gate kernel); the U kernel's last
// CTA posts flag_far, which the cluster polls before it loads block jb+2.  Flags only grow (host mirror `base`).
struct ChainArgs {
    float* A;        // element (panel row 0, panel column 0) of the packed FP32 master
    long lda;
    int D, pw;       // rows below the panel top, panel width (multiple of 16, D >= pw + 32)
    float* Yp;       // FP32 Y of the panel, element (panel row 0, panel column 0)
    long ldyp;
    float* Tslots;   // block jb's T at Tslots + jb * tstride (16 x 16, ld 16)
    int tstride;
    unsigned* flag_done;
    unsigned* flag_started;  // optional: base + 1 as soon as the cluster is resident (diagnostic: nothing waits on it)
    const unsigned* flag_far;
    unsigned base;
    float* srep[2];  // S replica pairs of the far updates (blocks alternate); the cluster clears rows 0..15 of both replicas of
                     // srep[jb & 1] (replica stride RMAX * SLD, row stride SLD) before it posts block jb
    unsigned wait0;  // block 0 waits for flag_far >= wait0 when have_wait0 (the previous panel's updates of THIS panel's columns)
    int have_wait0;
    long long* dbg;  // optional: 8 globaltimer stamps per block (CTA 0, thread 0), tools/chain_probe.py
};

__device__ __forceinline__ long long gtime_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CHAIN_STAMP(k)                                                       \
    if (a.dbg && crank == 0 && tid == 0) a.dbg[jb * 8 + (k)] = gtime_ns();


// Near update from shared memory.  ysl: this thread's Y_{jb-1} rows, float4 index (g * RPT + u) * NT + tid holds columns
// 4g..4g+3 of row u.  gtT: T_{jb-1} (upper triangular, zeros below).  pred: NW x 256 floats, pslotP: CSMAX x 256.
template <int RPT>
__device__ __forceinline__ void chain_near_update(float (&x)[RPT][16], const float4* ysl, const float (*gtT)[20], int tid, int lane,
                                                  int warp, int CS, unsigned crank, float* pred, float (*pslotP)[256], float* ptot,
                                                  float* ssm, uint64_t* mbarP, uint32_t parity) {
    constexpr int B = 16;
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(mbarP, (uint32_t)CS * 1024u);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        float acc[4][B];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < B; ++b) acc[k][b] = 0.f;
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const float4 y = ysl[(g * RPT + u) * NT + tid];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                acc[0][b] = fmaf(y.x, x[u][b], acc[0][b]);
                acc[1][b] = fmaf(y.y, x[u][b], acc[1][b]);
                acc[2][b] = fmaf(y.z, x[u][b], acc[2][b]);
                acc[3][b] = fmaf(y.w, x[u][b], acc[3][b]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            warp_transpose_reduce<B>(acc[k], lane);  // lane l: column l >> 1
            if ((lane & 1) == 0) pred[warp * 256 + (4 * g + k) * B + (lane >> 1)] = acc[k][0];
        }
    }
    __syncthreads();
    {
        float cs = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) cs += pred[w * 256 + tid];
        ptot[tid] = cs;  // this CTA's part of P, element (tid >> 4, tid & 15)
    }
    __syncthreads();
    if (CS > 1) {
        const uint32_t bar_local = smem_addr(mbarP);
        for (int o = tid; o < CS * 64; o += NT) {
            const unsigned peer = (unsigned)(o >> 6);
            const int ch = o & 63;
            const float4 v = *reinterpret_cast<const float4*>(&ptot[4 * ch]);
            st_async_v4(map_to_cta(smem_addr(&pslotP[crank][4 * ch]), peer), v, map_to_cta(bar_local, peer));
        }
        mbar_wait_cluster(mbarP, parity);
        float t = 0.f;
        for (int c = 0; c < CS; ++c) t += pslotP[c][tid];  // fixed order: every CTA gets the same bits
        __syncthreads();  // every sender has read its chunk of ptot
        ptot[tid] = t;
    }
    __syncthreads();
    {
        const int a2 = tid >> 4, b2 = tid & 15;
        float sv = 0.f;
#pragma unroll
        for (int c = 0; c < B; ++c) sv = fmaf(gtT[c][a2], ptot[c * B + b2], sv);  // S = T^T P
        ssm[tid] = sv;
    }
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        float sreg[4][B];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < B / 4; ++q) {
                const float4 s4 = *reinterpret_cast<const float4*>(&ssm[(4 * g + k) * B + 4 * q]);
                sreg[k][4 * q] = s4.x; sreg[k][4 * q + 1] = s4.y; sreg[k][4 * q + 2] = s4.z; sreg[k][4 * q + 3] = s4.w;
            }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const float4 y = ysl[(g * RPT + u) * NT + tid];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                float d = y.x * sreg[0][b];
                d = fmaf(y.y, sreg[1][b], d);
                d = fmaf(y.z, sreg[2][b], d);
                d = fmaf(y.w, sreg[3][b], d);
                x[u][b] -= d;
            }
        }
    }
}

template <int RPT>
__global__ void __launch_bounds__(NT, 1) panel_chain_kernel(ChainArgs a, int CS) {
    constexpr int B = 16;
    __shared__ __align__(16) float red[2][NW][B];
    __shared__ __align__(16) float prow[2][B];
    __shared__ __align__(16) float slot[2][CSMAX][B];
    __shared__ __align__(16) float pslot[2][B];
    __shared__ __align__(16) float tauS[NW][B];
    __shared__ __align__(16) float gt[B][B + 4];
    __shared__ float diag[B];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ __align__(16) uint4 tiles[NW][32 * (B / 4)];  // warp transposition tiles; the near update's per-warp partial P
    __shared__ __align__(16) float pslotP[CSMAX][256];
    __shared__ __align__(16) float fusedsm[2 * 256];
    __shared__ __align__(8) uint64_t mbarP;
    extern __shared__ __align__(16) float4 ysl[];  // 4 * RPT * NT float4: this thread's rows of the previous block's Y

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = (CS > 1) ? cluster_ctarank() : 0u;
    const int D = a.D;
    const int nblk = a.pw / B;
    const int rbase = (int)crank * (NT * RPT) + tid;
    const int wrow0 = rbase - lane;
    const long lda = a.lda, ldyp = a.ldyp;

    for (int idx = tid; idx < B * (B + 4); idx += NT) (&gt[0][0])[idx] = 0.f;
    if (CS > 1) {
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            mbar_init(&mbarP, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();
    }
    pdl_wait();
    if (a.flag_started && crank == 0 && tid == 0) atomicExch(a.flag_started, a.base + 1u);

    long long dummy_acc[1];
    long long dummy_prev = 0;
#pragma unroll 1
    for (int jb = 0; jb < nblk; ++jb) {
        const int roff = jb * B;
        float* Ab = a.A + roff;  // the block's first column
        CHAIN_STAMP(0);
        if (a.flag_far && (jb >= 2 || (jb == 0 && a.have_wait0))) {
            if (tid == 0) {
                const unsigned want = jb >= 2 ? a.base + (unsigned)(jb - 1) : a.wait0;  // far(jb-2) posted base + jb - 1
                while ((int)(ld_acquire_u32(a.flag_far) - want) < 0) __nanosleep(64);
            }
            __syncthreads();
        }
        CHAIN_STAMP(1);
        float x[RPT][B];
        {
            uint4 t[RPT][B / 4];
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int g0 = wrow0 + u * NT;
                warp_tile_fetch<B / 4>(reinterpret_cast<const char*>(Ab + (size_t)g0 * lda), (size_t)lda * 4, D - g0, t[u], lane);
            }
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                uint4 v[B / 4];
                warp_tile_transpose<B / 4>(t[u], v, tiles[warp], lane);
#pragma unroll
                for (int q = 0; q < B / 4; ++q) {
                    x[u][4 * q] = __uint_as_float(v[q].x); x[u][4 * q + 1] = __uint_as_float(v[q].y);
                    x[u][4 * q + 2] = __uint_as_float(v[q].z); x[u][4 * q + 3] = __uint_as_float(v[q].w);
                }
            }
        }
        __syncthreads();
        CHAIN_STAMP(2);
        if (jb > 0) {
            chain_near_update<RPT>(x, ysl, gt, tid, lane, warp, CS, crank, reinterpret_cast<float*>(&tiles[0][0]), pslotP, fusedsm,
                                   fusedsm + 256, &mbarP, (uint32_t)((jb - 1) & 1));
            // rows [roff-16, roff) of these columns are final R entries now
            if (crank == 0 && tid >= roff - B && tid < roff) {
#pragma unroll
                for (int q = 0; q < B / 4; ++q)
                    *reinterpret_cast<float4*>(Ab + (size_t)tid * lda + 4 * q) = make_float4(x[0][4 * q], x[0][4 * q + 1], x[0][4 * q + 2], x[0][4 * q + 3]);
            }
            __syncthreads();
        }
        CHAIN_STAMP(3);
        {
            StepMem<B> M{red, prow, slot, pslot, tauS, gt, diag, mbar};
            StepCtx sc{tid, lane, warp, CS, rbase, B, B, roff, crank, false, dummy_acc, &dummy_prev};
            factor_steps<B, RPT>(x, M, sc);
        }
        __syncthreads();
        CHAIN_STAMP(4);
        tinv_smem<B, B + 4>(gt, &red[0][0][0], tid);
        if (jb == nblk - 1) pdl_launch_dependents();
        CHAIN_STAMP(5);

        // the block's first 32 rows of the packed factor (CTA 0, u = 0); the rest is written per panel by panel_finalize_kernel
        if (crank == 0 && tid >= roff && tid < roff + 32) {
            const int i = tid, i2 = i - roff;
            if (i2 >= B) {
#pragma unroll
                for (int q = 0; q < B / 4; ++q)
                    *reinterpret_cast<float4*>(Ab + (size_t)(i + 1) * lda + 4 * q) = make_float4(x[0][4 * q], x[0][4 * q + 1], x[0][4 * q + 2], x[0][4 * q + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < B; ++c) Ab[(size_t)(i + (i2 >= c ? 1 : 0)) * lda + c] = x[0][c];
                Ab[(size_t)i * lda + i2] = diag[i2];
            }
        }
        // Y: zero above the block's diagonal (incl. all rows above the block)
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int i = rbase + u * NT;
            if (i < roff + B) {
#pragma unroll
                for (int c = 0; c < B; ++c)
                    if (i < roff + c) x[u][c] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int g0 = wrow0 + u * NT;
            rows_out_f32<B>(a.Yp + (size_t)g0 * ldyp + roff, ldyp, D - g0, x[u], tiles[warp], lane);
#pragma unroll
            for (int g = 0; g < 4; ++g) ysl[(g * RPT + u) * NT + tid] = make_float4(x[u][4 * g], x[u][4 * g + 1], x[u][4 * g + 2], x[u][4 * g + 3]);
        }
        if (crank == 0) {
            float* Tj = a.Tslots + (size_t)jb * a.tstride;
            const int t = tid >> 4, c = tid & 15;
            Tj[tid] = (t <= c) ? gt[t][c] : 0.f;
        }
        if (jb + 2 < nblk) {  // far(jb) accumulates into srep[jb & 1]; its previous user far(jb-2) was awaited above
            constexpr int Q = SLD / 4;  // float4 per row
            for (int idx = (int)crank * NT + tid; idx < 2 * B * Q; idx += CS * NT) {
                const int rep = idx / (B * Q), rem = idx - rep * (B * Q);
                reinterpret_cast<float4*>(a.srep[jb & 1] + (size_t)rep * RMAX * SLD + (size_t)(rem / Q) * SLD)[rem % Q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        CHAIN_STAMP(6);
        __threadfence();
        if (CS > 1) cluster_sync_all(); else __syncthreads();
        if (crank == 0 && tid == 0) atomicExch(a.flag_done, a.base + (unsigned)(jb + 1));   // (ordered behind every CTA's fenced stores by the barrier)
        CHAIN_STAMP(7);
    }
    // shared memory must stay alive until no peer can signal into it any more (the last cluster barrier above covers it)
}

__global__ void chain_gate_kernel(const unsigned* flag, unsigned want) {
    while ((int)(ld_acquire_u32(flag) - want) < 0) __nanosleep(200);
}

// ------------------------------------------------------------------ deferred outputs of a multi-block panel
// The register-block kernels touch their outputs as 64-byte (A, stride lda) and 32-byte (16-bit Y, stride ldh) row
// pieces through 16 SMs: ~8 of their 40 us at D = 32768.  With defer_out they only store the FP32 Y (compact, 512-byte
// rows, needed by the in-panel updates at once) and the top 32 rows of the block; this kernel then writes, once per
// panel and device-wide, the packed factor (w shifted one row down: A[i + 1][c] = Y[i][c] for rows below the top 32 of
// c's block) and the 16-bit Y (all rows, zeros above the panel included) from the FP32 Y.
//   Yp: D x ncols (ldyp) FP32 Y of the panel, zeros above each block's diagonal.   A: element (panel row 0, panel col 0).
//   Y16: element (row blk_row0 = panel row -zr, panel col 0) or null.   One thread per 4 columns of a row.
__global__ void __launch_bounds__(256) panel_finalize_kernel(const float* Yp, long ldyp, float* A, long lda, void* Y16, long ldy16,
                                                             int D, int ncols, int B, int zr, int bf16) {
    pdl_launch_dependents();
    pdl_wait();
    const int cq = ncols >> 2;  // float4 chunks per row
    const long total = (long)(D + zr) * cq;
    const long stride = (long)gridDim.x * blockDim.x;
    auto emit = [&](long idx, const float4& y) {
        const long rr = idx / cq;          // 0 .. D + zr - 1: row of the 16-bit output, starting at blk_row0
        const int c = (int)(idx - rr * cq) * 4;
        const long i = rr - zr;            // panel row
        if (Y16) {
            uint2 h = make_uint2(pack16(y.x, y.y, bf16), pack16(y.z, y.w, bf16));
            *reinterpret_cast<uint2*>((char*)Y16 + ((size_t)rr * ldy16 + c) * 2) = h;
        }
        if (i >= 0) {
            const int j0 = (c / B) * B;    // first row / column of c's register block
            if (i >= j0 + 32) *reinterpret_cast<float4*>(A + (size_t)(i + 1) * lda + c) = y;
        }
    };
    auto fetch = [&](long idx) {
        const long rr = idx / cq;
        const long i = rr - zr;
        const int c = (int)(idx - rr * cq) * 4;
        return (i >= 0) ? __ldcg(reinterpret_cast<const float4*>(Yp + (size_t)i * ldyp + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // four independent 16-byte loads in flight per thread: the pass sits on the panel chain between two cluster kernels and a
    // one-load-per-iteration loop was latency-bound ([B200] 25-33 us for 40 MB of traffic at D = 32768 on a 64-SM partition)
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + 3 * stride < total; idx += 4 * stride) {
        const float4 y0 = fetch(idx), y1 = fetch(idx + stride), y2 = fetch(idx + 2 * stride), y3 = fetch(idx + 3 * stride);
        emit(idx, y0); emit(idx + stride, y1); emit(idx + 2 * stride, y2); emit(idx + 3 * stride, y3);
    }
    for (; idx < total; idx += stride) emit(idx, fetch(idx));
}

// ------------------------------------------------------------------ level 1: in-panel update
// S'[rep][t][c] += sum_rows Y[row][t] * A[row][c]   (FP32; B x ncols, rows split over the grid)
// 512 threads = 4 row groups x 128 columns; RB rows in flight per thread (memory-level parallelism
// is what bounds these skinny passes: 512 x RB x 4 B in flight per SM).
// rows in flight per thread: the loops are latency-bound (one round trip per batch), so a CTA's rows
// should be covered in about two batches
template <int B> struct SuRb { static constexpr int value = (B == 16) ? 32 : 16; };
template <int B>
__global__ void __launch_bounds__(512) inpanel_s_kernel(const float* __restrict__ W, long ldw, const float* __restrict__ A,
                                                         long lda, int D, int ncols, float* __restrict__ Srep, int rows_per_cta) {
    constexpr int SU_RB = SuRb<B>::value;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, c = tid & 127, rg = tid >> 7;
    const int r0 = blockIdx.x * rows_per_cta;
    int nrows = D - r0;
    if (nrows > rows_per_cta) nrows = rows_per_cta;
    if (nrows <= 0) return;
    const int c0 = blockIdx.y * 128;
    const bool on = (c0 + c) < ncols;
    const float* Ac = A + (size_t)r0 * lda + c0 + c;
    pdl_launch_dependents();
    pdl_wait();
    // first batch of A goes out before the Y chunk is staged (overlaps the two global latencies)
    float av[SU_RB];
#pragma unroll
    for (int u = 0; u < SU_RB; ++u) {
        const int r2 = rg + 4 * u;
        av[u] = (on && r2 < nrows) ? __ldcg(Ac + (size_t)r2 * lda) : 0.f;
    }
    for (int idx = tid; idx < nrows * B; idx += 512) {
        int rr = idx / B, t = idx - rr * B;
        sm[idx] = __ldcg(&W[(size_t)(r0 + rr) * ldw + t]);
    }
    __syncthreads();
    float acc[B];
#pragma unroll
    for (int t = 0; t < B; ++t) acc[t] = 0.f;
    for (int rr = rg; rr < nrows; rr += 4 * SU_RB) {
        float nx[SU_RB];
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) {  // next batch in flight while this one is consumed
            const int r2 = rr + 4 * SU_RB + 4 * u;
            nx[u] = (on && r2 < nrows) ? __ldcg(Ac + (size_t)r2 * lda) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) {
            const int r2 = rr + 4 * u;
            if (r2 < nrows) {
                const float* wr = sm + r2 * B;
#pragma unroll
                for (int q = 0; q < B / 4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * q);
                    acc[4 * q] = fmaf(w4.x, av[u], acc[4 * q]);
                    acc[4 * q + 1] = fmaf(w4.y, av[u], acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(w4.z, av[u], acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(w4.w, av[u], acc[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) av[u] = nx[u];
    }
    __syncthreads();  // Y chunk no longer needed: reuse shared memory for the row-group reduction
    if (rg > 0) {
#pragma unroll
        for (int t = 0; t < B; ++t) sm[((rg - 1) * B + t) * 128 + c] = acc[t];
    }
    __syncthreads();
    if (rg == 0 && on) {
        float* S = Srep + (size_t)(blockIdx.x & 1) * RMAX * SLD;
#pragma unroll
        for (int t = 0; t < B; ++t) {
            float v = acc[t] + sm[(0 * B + t) * 128 + c] + sm[(1 * B + t) * 128 + c] + sm[(2 * B + t) * 128 + c];
            atomicAdd(&S[t * SLD + c0 + c], v);
        }
    }
}

// A[row][c] -= sum_t Y[row][t] * S[t][c],  S = T^T (sum of the replicas)   [Q_j^T A = A - Y T^T (Y^T A)]
template <int B>
__global__ void __launch_bounds__(512) inpanel_u_kernel(const float* __restrict__ Y, long ldy, float* __restrict__ A, long lda,
                                                         int D, int ncols, const float* __restrict__ Srep,
                                                         const float* __restrict__ Tj, int rows_per_cta) {
    constexpr int SU_RB = SuRb<B>::value;
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(16) float Ts[B][B + 4];
    __shared__ float Sp[B][128];
    __shared__ float S2[B][128];
    const int tid = threadIdx.x, c = tid & 127, rg = tid >> 7;
    const int r0 = blockIdx.x * rows_per_cta;
    int nrows = D - r0;
    if (nrows > rows_per_cta) nrows = rows_per_cta;
    if (nrows <= 0) return;
    const int c0 = blockIdx.y * 128;
    const bool on = (c0 + c) < ncols;
    float* Ac = A + (size_t)r0 * lda + c0 + c;
    pdl_launch_dependents();
    pdl_wait();
    float av[SU_RB];
#pragma unroll
    for (int u = 0; u < SU_RB; ++u) {  // first batch of A in flight during the prologue
        const int r2 = rg + 4 * u;
        av[u] = (on && r2 < nrows) ? __ldcg(&Ac[(size_t)r2 * lda]) : 0.f;
    }
    // S' = sum of the replicas (B x 128 chunk), once per CTA
    for (int idx = tid; idx < B * 128; idx += 512) {
        const int t = idx >> 7, cc = idx & 127;
        float v = 0.f;
        if (c0 + cc < ncols) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) v += __ldcg(&Srep[(size_t)rep * RMAX * SLD + t * SLD + c0 + cc]);
        }
        Sp[t][cc] = v;
    }
    for (int idx = tid; idx < B * B; idx += 512) Ts[idx / B][idx % B] = __ldcg(&Tj[idx]);
    for (int idx = tid; idx < nrows * B; idx += 512) {
        int rr = idx / B, t = idx - rr * B;
        sm[idx] = __ldcg(&Y[(size_t)(r0 + rr) * ldy + t]);
    }
    __syncthreads();
    // S = T^T S': row group rg computes B/4 consecutive rows t of its column; T is zero below its
    // diagonal, so the sum runs over all u2 with 16-byte broadcast reads of T (shared-memory bandwidth bound)
    {
        constexpr int TQ = B / 4;
        float v[TQ];
#pragma unroll
        for (int q = 0; q < TQ; ++q) v[q] = 0.f;
#pragma unroll 4
        for (int u2 = 0; u2 < B; ++u2) {
            const float sp = Sp[u2][c];
#pragma unroll
            for (int q = 0; q < TQ; q += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(&Ts[u2][rg * TQ + q]);
                v[q] = fmaf(t4.x, sp, v[q]);
                v[q + 1] = fmaf(t4.y, sp, v[q + 1]);
                v[q + 2] = fmaf(t4.z, sp, v[q + 2]);
                v[q + 3] = fmaf(t4.w, sp, v[q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < TQ; ++q) S2[rg * TQ + q][c] = v[q];
    }
    __syncthreads();
    if (!on) return;
    float s[B];
#pragma unroll
    for (int t = 0; t < B; ++t) s[t] = S2[t][c];
    for (int rr = rg; rr < nrows; rr += 4 * SU_RB) {
        float nx[SU_RB];
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) {
            const int r2 = rr + 4 * SU_RB + 4 * u;
            nx[u] = (r2 < nrows) ? __ldcg(&Ac[(size_t)r2 * lda]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) {
            const int r2 = rr + 4 * u;
            if (r2 < nrows) {
                const float* yr = sm + r2 * B;
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int q = 0; q < B / 4; ++q) {
                    const float4 y4 = *reinterpret_cast<const float4*>(yr + 4 * q);
                    d0 = fmaf(y4.x, s[4 * q], d0);
                    d1 = fmaf(y4.y, s[4 * q + 1], d1);
                    d0 = fmaf(y4.z, s[4 * q + 2], d0);
                    d1 = fmaf(y4.w, s[4 * q + 3], d1);
                }
                Ac[(size_t)r2 * lda] = av[u] - (d0 + d1);
            }
        }
#pragma unroll
        for (int u = 0; u < SU_RB; ++u) av[u] = nx[u];
    }
}

// ---- vectorised in-panel kernels (16-byte aligned A_rest, ncols % 4 == 0): one thread owns FOUR
// columns and all B reflectors, i.e. 4*B FMAs per 16-byte load and B/4 broadcast LDS.128 (the
// one-column kernels above issue ~5x more instructions per FMA and are kept for unaligned shapes).
// The last CTA to finish (ticket counter) folds the replicas and applies T^T once, so that the
// update kernel starts from the final S = T^T (Y^T A_rest).
template <int B> struct Su4 {
    static constexpr int NTHR = (B == 16) ? 512 : 256;
    static constexpr int RB = 8;  // rows in flight per thread
};

template <int B>
__global__ void __launch_bounds__(Su4<B>::NTHR) inpanel_s4_kernel(const float* __restrict__ Y, long ldy, const float* __restrict__ A,
                                                                  long lda, int D, int ncols, float* __restrict__ Srep,
                                                                  const float* __restrict__ Tj,
                                                                  float* __restrict__ Sfin, int rows_per_cta, int ysm_floats) {
    constexpr int NTHR = Su4<B>::NTHR, NWARP = NTHR / 32, RB = Su4<B>::RB;
    extern __shared__ __align__(16) float sm[];
    float* ysm = sm;               // rows_per_cta x B  (later: T, B x (B+4))
    float* red = sm + ysm_floats;  // B x 128           (later: S')
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * rows_per_cta;
    int nrows = D - r0;
    if (nrows > rows_per_cta) nrows = rows_per_cta;
    const int c0 = blockIdx.y * 128, col = c0 + 4 * lane;
    const bool on = col < ncols;
    pdl_wait();
    pdl_launch_dependents();   // (after the wait: the U kernel behind this one fetches A_rest and Y before ITS wait)
    const float* Ap = A + (size_t)r0 * lda + col;
    float4 a4[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {  // first batch in flight while Y is staged
        const int r2 = warp + u * NWARP;
        a4[u] = (on && r2 < nrows) ? __ldcg(reinterpret_cast<const float4*>(Ap + (size_t)r2 * lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int idx = tid; idx < nrows * B; idx += NTHR) {
        const int rr = idx / B, t = idx - rr * B;
        ysm[idx] = __ldcg(&Y[(size_t)(r0 + rr) * ldy + t]);
    }
    __syncthreads();
    float acc[B][4];
#pragma unroll
    for (int t = 0; t < B; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    for (int rb = warp; rb < nrows; rb += NWARP * RB) {
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r2 = rb + u * NWARP;
            if (r2 < nrows) {
                const float* yr = ysm + r2 * B;
#pragma unroll
                for (int q = 0; q < B / 4; ++q) {
                    const float4 y4 = *reinterpret_cast<const float4*>(yr + 4 * q);
                    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[4 * q + k][0] = fmaf(yv[k], a4[u].x, acc[4 * q + k][0]);
                        acc[4 * q + k][1] = fmaf(yv[k], a4[u].y, acc[4 * q + k][1]);
                        acc[4 * q + k][2] = fmaf(yv[k], a4[u].z, acc[4 * q + k][2]);
                        acc[4 * q + k][3] = fmaf(yv[k], a4[u].w, acc[4 * q + k][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r2 = rb + NWARP * RB + u * NWARP;
            a4[u] = (on && r2 < nrows) ? __ldcg(reinterpret_cast<const float4*>(Ap + (size_t)r2 * lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // cross-warp tree through shared memory (conflict-free 16-byte slots), then warp 0 adds the
    // CTA's partial S' into its replica
    __syncthreads();  // Y chunk no longer needed
    for (int half = NWARP / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
            float4* dst = reinterpret_cast<float4*>(sm) + ((size_t)(warp - half) * B) * 32 + lane;
#pragma unroll
            for (int t = 0; t < B; ++t) dst[t * 32] = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        }
        __syncthreads();
        if (warp < half) {
            const float4* src = reinterpret_cast<const float4*>(sm) + ((size_t)warp * B) * 32 + lane;
#pragma unroll
            for (int t = 0; t < B; ++t) {
                const float4 v = src[t * 32];
                acc[t][0] += v.x; acc[t][1] += v.y; acc[t][2] += v.z; acc[t][3] += v.w;
            }
        }
        __syncthreads();
    }
    // T^T is linear: every CTA applies it to its OWN partial S' and adds the result into one of two replicas of the
    // final S (no ticket, no last-CTA fold: four dependent global round trips fewer on the panel chain)
    float* Ts2 = ysm;  // B x (B + 4); the tree above is done with this memory
    if (warp == 0) {
#pragma unroll
        for (int t = 0; t < B; ++t) *reinterpret_cast<float4*>(&red[t * 128 + 4 * lane]) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    }
    for (int idx = tid; idx < B * B; idx += NTHR) Ts2[(idx / B) * (B + 4) + (idx % B)] = __ldcg(&Tj[idx]);
    __syncthreads();
    constexpr int NG = NTHR / 128, TQ = B / NG;  // thread: column cc, TQ consecutive rows t of S
    const int cc = tid & 127, tg = tid >> 7;
    float v[TQ];
#pragma unroll
    for (int q = 0; q < TQ; ++q) v[q] = 0.f;
#pragma unroll 4
    for (int u2 = 0; u2 < B; ++u2) {  // T is zero below its diagonal
        const float sp = red[u2 * 128 + cc];
#pragma unroll
        for (int q = 0; q < TQ; q += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(&Ts2[u2 * (B + 4) + tg * TQ + q]);
            v[q] = fmaf(t4.x, sp, v[q]);
            v[q + 1] = fmaf(t4.y, sp, v[q + 1]);
            v[q + 2] = fmaf(t4.z, sp, v[q + 2]);
            v[q + 3] = fmaf(t4.w, sp, v[q + 3]);
        }
    }
    if (c0 + cc < ncols) {
        float* S = Srep + (size_t)(blockIdx.x & 1) * RMAX * SLD + c0 + cc;
#pragma unroll
        for (int q = 0; q < TQ; ++q) atomicAdd(&S[(size_t)(tg * TQ + q) * SLD], v[q]);
    }
}

template <int B>
__global__ void __launch_bounds__(Su4<B>::NTHR) inpanel_u4_kernel(const float* __restrict__ Y, long ldy, float* __restrict__ A, long lda,
                                                                  int D, int ncols, const float* __restrict__ Sfin, int nrep, int rows_per_cta,
                                                                  unsigned* post, unsigned post_val, unsigned* ticket) {
    // Sfin: nrep replicas (stride RMAX * SLD) of S = T^T (Y^T A_rest), summed here
    // post != null (persistent panel chain): the last CTA to finish publishes post_val (ticket counter, reset for the next launch)
    constexpr int NTHR = Su4<B>::NTHR, NWARP = NTHR / 32, RB = Su4<B>::RB;
    extern __shared__ __align__(16) float sm[];
    float* ysm = sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * rows_per_cta;
    int nrows = D - r0;
    if (nrows > rows_per_cta) nrows = rows_per_cta;
    const int c0 = blockIdx.y * 128, col = c0 + 4 * lane;
    const bool on = col < ncols;
    pdl_launch_dependents();
    // This launch always follows the S kernel of the same update directly (launch_su).  That kernel only READS A_rest and Y,
    // and whatever wrote them had finished before its CTAs passed their own griddepcontrol.wait -- which all of them had done
    // when this grid was let in.  The first batch of rows and the Y slab are therefore fetched BEFORE the wait; only S needs it.
    float* Ap = A + (size_t)r0 * lda + col;
    float4 a4[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        const int r2 = warp + u * NWARP;
        a4[u] = (on && r2 < nrows) ? __ldcg(reinterpret_cast<const float4*>(Ap + (size_t)r2 * lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int idx = tid; idx < nrows * B; idx += NTHR) {
        const int rr = idx / B, t = idx - rr * B;
        ysm[idx] = __ldcg(&Y[(size_t)(r0 + rr) * ldy + t]);
    }
    pdl_wait();
    float sv[B][4];
#pragma unroll
    for (int t = 0; t < B; ++t) {
        float4 s4 = on ? __ldcg(reinterpret_cast<const float4*>(Sfin + (size_t)t * SLD + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (on && nrep > 1) {
            const float4 s5 = __ldcg(reinterpret_cast<const float4*>(Sfin + (size_t)RMAX * SLD + (size_t)t * SLD + col));
            s4.x += s5.x; s4.y += s5.y; s4.z += s5.z; s4.w += s5.w;
        }
        sv[t][0] = s4.x; sv[t][1] = s4.y; sv[t][2] = s4.z; sv[t][3] = s4.w;
    }
    __syncthreads();
    for (int rb = warp; on && rb < nrows; rb += NWARP * RB) {
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r2 = rb + u * NWARP;
            if (r2 < nrows) {
                const float* yr = ysm + r2 * B;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < B / 4; ++q) {
                    const float4 y4 = *reinterpret_cast<const float4*>(yr + 4 * q);
                    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        d[0] = fmaf(yv[k], sv[4 * q + k][0], d[0]);
                        d[1] = fmaf(yv[k], sv[4 * q + k][1], d[1]);
                        d[2] = fmaf(yv[k], sv[4 * q + k][2], d[2]);
                        d[3] = fmaf(yv[k], sv[4 * q + k][3], d[3]);
                    }
                }
                *reinterpret_cast<float4*>(Ap + (size_t)r2 * lda) = make_float4(a4[u].x - d[0], a4[u].y - d[1], a4[u].z - d[2], a4[u].w - d[3]);
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r2 = rb + NWARP * RB + u * NWARP;
            a4[u] = (r2 < nrows) ? __ldcg(reinterpret_cast<const float4*>(Ap + (size_t)r2 * lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (post) {
        __threadfence();   // this thread's stores before the CTA's ticket
        __syncthreads();
        if (tid == 0) {
            const unsigned total = gridDim.x * gridDim.y;
            if (atomicAdd(ticket, 1u) == total - 1) {
                *ticket = 0u;          // (no CTA of this launch touches it again; the next launch is stream-ordered behind)
                __threadfence();
                atomicExch(post, post_val);
            }
        }
    }
}

// ------------------------------------------------------------------ level 2: T = (striu(G) + I/2)^-1
// One CTA, recursive doubling: the 16 x 16 diagonal blocks are inverted column by column, then
// T12 = -T11 (G12 T22) merges pairs of blocks (16 -> 32 -> 64 -> 128).  G is read from global
// (ldg); T goes to T32 (pw x pw, zero below the diagonal) and optionally to a 16-bit copy.
constexpr int TLD = RMAX + 1;
__global__ void __launch_bounds__(1024) tinv_kernel(const float* __restrict__ G, long ldg, int pw, float* __restrict__ T32, int ldt,
                                                    void* __restrict__ T16, long ldt16, int bf16, const float* Sy, long ldsy, int ns,
                                                    void* __restrict__ S16, long lds16) {
    // Sy != null: also S = T^T Sy (pw x ns, ns <= RMAX; Sy FP32 with ld ldsy) -> S16 (16-bit, ld lds16)
    extern __shared__ __align__(16) float sm[];
    float* Ts = sm;               // RMAX x TLD
    float* Xs = sm + RMAX * TLD;  // 64 x 65 temp
    float* Ss = Xs + 64 * 65;     // pw x RMAX: Sy (only when Sy != null)
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (Sy) {
        for (int idx = tid; idx < pw * RMAX; idx += 1024) {
            const int t = idx / RMAX, c = idx - t * RMAX;
            Ss[idx] = (c < ns) ? __ldcg(&Sy[(size_t)t * ldsy + c]) : 0.f;
        }
    }
    int R = 16;
    while (R < pw) R *= 2;
    for (int idx = tid; idx < R * R; idx += 1024) {
        int t = idx / R, c = idx - t * R;
        Ts[t * TLD + c] = (t < c && c < pw) ? __ldcg(&G[(size_t)t * ldg + c]) : 0.f;
    }
    __syncthreads();
    // diagonal blocks: thread (block b, column c) runs the larft recurrence restricted to the block
    // column: x_c = 2, x_t = -2 * sum_{u=t+1..c} G[t][u] x_u  (back substitution of (striu(G)+I/2) x = e_c)
    float xv[16];
    const int b0 = tid & ~15, cc = tid & 15;
    if (tid < R) {
#pragma unroll
        for (int t = 15; t >= 0; --t) {
            float sacc = 0.f;
#pragma unroll
            for (int u2 = t + 1; u2 < 16; ++u2)
                if (u2 <= cc) sacc = fmaf(Ts[(b0 + t) * TLD + b0 + u2], xv[u2], sacc);
            xv[t] = (t == cc) ? 2.f : ((t < cc) ? -2.f * sacc : 0.f);
        }
    }
    __syncthreads();  // every column has read G before anyone overwrites it
    if (tid < R) {
#pragma unroll
        for (int t = 0; t < 16; ++t)
            if (t <= cc) Ts[(b0 + t) * TLD + b0 + cc] = xv[t];
    }
    __syncthreads();
    for (int h = 16; h < R; h *= 2) {
        // pairs: blocks [o, o+h) and [o+h, o+2h).  Each thread owns a 4 x 4 tile of the h x h products
        // (shared memory bandwidth bounds this kernel: 0.5 LDS per FMA instead of 2).
        const int npairs = R / (2 * h);
        const int tpr = h / 4;                 // tiles per row/column of one product
        const int ntiles = npairs * tpr * tpr;  // <= (R/2h) * h*h/16 = R*h/32 <= 256
        const bool on = tid < ntiles;
        const int pr = tid / (tpr * tpr), e = tid - pr * tpr * tpr;
        const int i0 = 4 * (e / tpr), j0 = 4 * (e % tpr), o = pr * 2 * h;
        float acc[4][4];
        if (on) {
            // X = G12 * T22 (T22 zero below its diagonal: fixed trip count)
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = 0.f;
            for (int u2 = 0; u2 < h; ++u2) {
                float ga[4], tb[4];
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2) ga[a2] = Ts[(o + i0 + a2) * TLD + o + h + u2];
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) tb[b2] = Ts[(o + h + u2) * TLD + o + h + j0 + b2];
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = fmaf(ga[a2], tb[b2], acc[a2][b2]);
            }
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) Xs[(pr * h + i0 + a2) * 65 + j0 + b2] = acc[a2][b2];  // pr*h + i < 64
        }
        __syncthreads();
        if (on) {
            // T12 = -T11 * X (T11 zero below its diagonal)
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = 0.f;
            for (int u2 = 0; u2 < h; ++u2) {
                float ta[4], xb[4];
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2) ta[a2] = Ts[(o + i0 + a2) * TLD + o + u2];
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) xb[b2] = Xs[(pr * h + u2) * 65 + j0 + b2];
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = fmaf(ta[a2], xb[b2], acc[a2][b2]);
            }
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) Ts[(o + i0 + a2) * TLD + o + h + j0 + b2] = -acc[a2][b2];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < pw * pw; idx += 1024) {
        int t = idx / pw, c = idx - t * pw;
        const float v = (t <= c) ? Ts[t * TLD + c] : 0.f;
        if (T32) T32[(size_t)t * ldt + c] = v;
        if (T16) store16(T16, (long)t * ldt16 + c, v, bf16);
    }
    if (Sy) {
        // S[a][c] = sum_{b <= a} T[b][a] Sy[b][c]: thread = 4 rows a x 4 columns c (one LDS.128 of Sy and four broadcast
        // reads of T per 16 FMAs; the first version, 16 rows x 1 column per thread, was bound by 2048 LDS per thread: +17 us)
        const int c4 = 4 * (tid & 31), a4 = 4 * (tid >> 5);
        if (a4 < pw) {
            float v[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.f;
            const int bmax = (a4 + 4 < pw) ? a4 + 4 : pw;   // T is zero below its diagonal
#pragma unroll 4
            for (int b = 0; b < bmax; ++b) {
                const float4 sy = *reinterpret_cast<const float4*>(&Ss[b * RMAX + c4]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float t = Ts[b * TLD + a4 + q];
                    v[q][0] = fmaf(t, sy.x, v[q][0]); v[q][1] = fmaf(t, sy.y, v[q][1]);
                    v[q][2] = fmaf(t, sy.z, v[q][2]); v[q][3] = fmaf(t, sy.w, v[q][3]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (a4 + q < pw) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c4 + k < ns) store16(S16, (long)(a4 + q) * lds16 + c4 + k, v[q][k], bf16);
                }
        }
    }
}

// ------------------------------------------------------------------ host side
struct ClusterCaps {
    int max_cs;
};
template <int B, int RPT>
int prepare_kernel(int* max_cs) {
    static int cached = 0;
    if (!cached) {
        cudaError_t e = cudaFuncSetAttribute(panel_block_kernel<B, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int mc = 8;
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(NT);
            cfg.dynamicSmemBytes = 0;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 16;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, panel_block_kernel<B, RPT>, &cfg) == cudaSuccess && n > 0) mc = 16;
            else cudaGetLastError();
        } else {
            cudaGetLastError();
        }
        cached = mc;
    }
    *max_cs = cached;
    return MPQR_OK;
}

template <int B, int RPT>
int launch_block_t(const BlockArgs& a, int CS, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CS > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CS;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    at[na++] = pdl_attr();
    cfg.attrs = at;
    cfg.numAttrs = na;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, panel_block_kernel<B, RPT>, a, CS));
    return MPQR_OK;
}

int g_max_cs = 0;
int max_cluster() {
    if (!g_max_cs) {
        int v[7];
        prepare_kernel<16, 1>(&v[0]);
        prepare_kernel<16, 2>(&v[1]);
        prepare_kernel<16, 4>(&v[2]);
        prepare_kernel<16, 8>(&v[3]);
        prepare_kernel<32, 1>(&v[4]);
        prepare_kernel<32, 2>(&v[5]);
        prepare_kernel<32, 4>(&v[6]);
        int mc = v[0];
        for (int i = 1; i < 7; ++i)
            if (v[i] < mc) mc = v[i];
        g_max_cs = mc;
    }
    return g_max_cs;
}

// rows a block kernel of width B can hold
long block_capacity(int B) { return (long)max_cluster() * NT * (B == 32 ? 4 : 8); }

// Picks (RPT, CS) for D rows; returns false if the block does not fit one cluster.
bool pick_shape(int B, int D, int force_cs, int force_rpt, int* rpt, int* cs) {
    const int mc = max_cluster();
    const int max_rpt = (B == 32) ? 4 : 8;
    if ((long)D > (long)mc * NT * max_rpt) return false;
    int R = 1, C = 1;
    if (D <= NT) { R = 1; C = 1; }
    else if (D <= 2 * NT) { R = 2; C = 1; }
    else if (D <= 4 * NT) { R = 4; C = 1; }
    else {
        // prefer one row per thread and more CTAs: the exchange costs the same for any CS > 1
        R = 1;
        C = 2;
        while (C < mc && (long)C * NT * R < D) C *= 2;
        while ((long)C * NT * R < D) R *= 2;
    }
    if (force_rpt > 0 && force_rpt <= max_rpt) {
        R = force_rpt;
        C = 1;
        while ((long)C * NT * R < D && C < mc) C *= 2;
        if ((long)C * NT * R < D) return false;
    }
    if (force_cs > 0 && force_cs <= mc) {
        C = force_cs;
        if (force_rpt <= 0) R = 1;
        while ((long)C * NT * R < D && R < max_rpt) R *= 2;
        if ((long)C * NT * R < D) return false;
    }
    *rpt = R;
    *cs = C;
    return true;
}

int launch_block(int B, const BlockArgs& a, int RPT, int CS, cudaStream_t st) {
    if (B == 32) {
        if (RPT == 1) return launch_block_t<32, 1>(a, CS, st);
        if (RPT == 2) return launch_block_t<32, 2>(a, CS, st);
        return launch_block_t<32, 4>(a, CS, st);
    }
    if (RPT == 1) return launch_block_t<16, 1>(a, CS, st);
    if (RPT == 2) return launch_block_t<16, 2>(a, CS, st);
    if (RPT == 4) return launch_block_t<16, 4>(a, CS, st);
    return launch_block_t<16, 8>(a, CS, st);
}

// workspace layout (floats): Y32p [2 x rows x RMAX] | Tsl [2 x 8 x 256] | Wj [rows x 32] | Srep [NREP x RMAX x SLD] | G [RMAX x RMAX] |
// T32 [RMAX x RMAX] | T16 [RMAX x RMAX 16-bit]
struct Ws {
    float* Y32p;     // FP32 Y of the panel; two buffers: the chain kernel of panel p+1 writes one while panel p's finalize /
    float* Y32p2;    // side updates still read the other
    float* Tsl;      // block T slots of the chain flow: 2 buffers x 8 blocks x 256 floats
    float* Wj;
    float* Srep;     // NREP replicas | 4 floats (ticket counter) : cleared together by the block kernel
    float* Sfin;     // RMAX x SLD: S = T^T (Y^T A_rest)
    float* G;
    float* T32;
    void* T16;
};
Ws carve(float* ws, long rows) {
    Ws w;
    w.Y32p = ws;
    w.Y32p2 = w.Y32p + (size_t)rows * RMAX;
    w.Tsl = w.Y32p2 + (size_t)rows * RMAX;
    w.Wj = w.Tsl + 2 * 8 * 256;
    w.Srep = w.Wj + (size_t)rows * 32;
    w.Sfin = w.Srep + (size_t)NREP * RMAX * SLD + 4 + 256;  // (+ ticket counter + 16 x 16 cross-Gram accumulator)
    w.G = w.Sfin + (size_t)RMAX * SLD;              // RMAX x 2 RMAX: [G | Sy] of the merged Gram / next-panel product
    w.T32 = w.G + (size_t)2 * RMAX * RMAX;
    w.T16 = (void*)(w.T32 + (size_t)RMAX * RMAX);   // two buffers (consecutive panels alternate: a deferred W = Y T reads one
    return w;                                       // while the next panel's T kernel writes the other)
}

int launch_tinv(const float* G, long ldg, int pw, float* T32, int ldt, void* T16, long ldt16, int bf16, size_t smem, cudaStream_t st,
                const float* Sy = nullptr, long ldsy = 0, int ns = 0, void* S16 = nullptr, long lds16 = 0) {
    cudaLaunchAttribute pat[1] = {pdl_attr()};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.stream = st; cfg.attrs = pat; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = smem;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, tinv_kernel, G, ldg, pw, T32, ldt, T16, ldt16, bf16, Sy, ldsy, ns, S16, lds16));
    return MPQR_OK;
}

template <int B>
int su_attrs() {
    MPQR_TRY(func_attr_once((const void*)inpanel_s_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    MPQR_TRY(func_attr_once((const void*)inpanel_u_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MPQR_TRY(func_attr_once((const void*)inpanel_s4_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MPQR_TRY(func_attr_once((const void*)inpanel_u4_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    return MPQR_OK;
}

template <int B>
int launch_su(const float* Tj, const float* Yj, long ldy, float* Arest, long lda, int D, int ncols, float* Srep, float* Sfin,
              int num_sms, cudaStream_t st, long* launches, const ProfHook* prof, bool pdl_first = true,
              unsigned* post = nullptr, unsigned post_val = 0, unsigned* ticket = nullptr, int max_rows = 512) {
    // one wave of CTAs over the SMs this stream may use (up to max_rows rows = 32-64 KB of staged Y per CTA);
    // taller blocks take k balanced waves
    const int waves = ceil_div(D, max_rows * num_sms);
    int rows = ceil_div(D, waves * num_sms);
    rows = round_up(rows < 16 ? 16 : rows, 16);
    if (rows > max_rows) rows = max_rows;
    MPQR_TRY(su_attrs<B>());
    dim3 grid(ceil_div(D, rows), ceil_div(ncols, 128));
    cudaLaunchAttribute pat[1] = {pdl_attr()};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.stream = st; cfg.attrs = pat; cfg.numAttrs = 1;
    const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(Arest) & 15) == 0) && ((ncols & 3) == 0) && ncols <= SLD;
    if (!vec && post) {   // only the vectorised pair knows the chain's flags (panel_chain_ok guarantees its alignment)
        set_error("in-panel update: flag-ordered launch needs the vectorised kernels (lda=%ld ncols=%d)", lda, ncols);
        return MPQR_EINVAL;
    }
    if (prof) prof->begin(prof->ctx, 5, st, 2.0 * D * ncols * B, 4.0 * D * (ncols + B));
    if (vec) {
        int ysm_floats = rows * B;
        if (ysm_floats < B * (B + 4)) ysm_floats = B * (B + 4);
        cfg.blockDim = dim3(Su4<B>::NTHR);
        size_t s4_floats = (size_t)ysm_floats + B * 128 + 512;  // (+ scratch of the T_32 assembly)
        const size_t tree_floats = (size_t)(Su4<B>::NTHR / 32) * B * 64;
        if (s4_floats < tree_floats) s4_floats = tree_floats;
        cfg.dynamicSmemBytes = s4_floats * sizeof(float);
        // pdl_first = false: the S kernel must not become resident (and hold its SMs idle) while the register-block
        // kernel before it is still running: those SMs belong to the side stream's updates during that time
        if (!pdl_first) cfg.numAttrs = 0;
        MPQR_CUDA(cudaLaunchKernelEx(&cfg, inpanel_s4_kernel<B>, Yj, ldy, (const float*)Arest, lda, D, ncols, Srep, Tj, Sfin, rows,
                                     ysm_floats));
        cfg.numAttrs = 1;
        if (prof) { prof->end(prof->ctx, st); prof->begin(prof->ctx, 7, st, 2.0 * D * ncols * B, 4.0 * D * (2 * ncols + B)); }
        cfg.dynamicSmemBytes = (size_t)rows * B * sizeof(float);
        MPQR_CUDA(cudaLaunchKernelEx(&cfg, inpanel_u4_kernel<B>, Yj, ldy, Arest, lda, D, ncols, (const float*)Srep, 2, rows, post, post_val, ticket));
    } else {
        cfg.blockDim = dim3(512);
        cfg.dynamicSmemBytes = (size_t)((rows * B > 3 * B * 128) ? rows * B : 3 * B * 128) * sizeof(float);
        MPQR_CUDA(cudaLaunchKernelEx(&cfg, inpanel_s_kernel<B>, Yj, ldy, (const float*)Arest, lda, D, ncols, Srep, rows));
        if (prof) { prof->end(prof->ctx, st); prof->begin(prof->ctx, 7, st, 2.0 * D * ncols * B, 4.0 * D * (2 * ncols + B)); }
        cfg.dynamicSmemBytes = (size_t)rows * B * sizeof(float);
        MPQR_CUDA(cudaLaunchKernelEx(&cfg, inpanel_u_kernel<B>, Yj, ldy, Arest, lda, D, ncols, (const float*)Srep, Tj, rows));
    }
    if (prof) prof->end(prof->ctx, st);
    if (launches) *launches += 2;
    return MPQR_OK;
}

// ---- persistent panel chain: host side
// Stream memory operation (driver API, resolved at run time like the green-context calls in api.cu): the side stream waits
// on the kernel's progress flag without any kernel of ours occupying an SM.  MPQR_GATE_KERNEL=1 (or a driver without the
// entry point) uses a one-thread gate kernel instead.
struct MemopApi {
    CUresult (*Wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    bool ok;
};
const MemopApi* memop_api() {
    static MemopApi api{};
    static bool tried = false;
    if (!tried) {
        tried = true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        api.ok = get("cuStreamWaitValue32", (void**)&api.Wait32);
        if (!api.ok) cudaGetLastError();
    }
    return &api;
}
int stream_wait_geq(cudaStream_t st, unsigned* flag, unsigned want) {
    const MemopApi* m = memop_api();
    if (m->ok && !getenv("MPQR_GATE_KERNEL")) {   // (read per call: the panel tests switch it inside one process)
        if (m->Wait32((CUstream)st, (CUdeviceptr)(uintptr_t)flag, want, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS) return MPQR_OK;
        set_error("cuStreamWaitValue32 failed");
        return MPQR_ECUDA;
    }
    chain_gate_kernel<<<1, 1, 0, st>>>(flag, want);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}
// CUDA loads kernels lazily, and loading one may need a context synchronisation: a first-time load of a side-stream
// kernel WHILE the cluster spins on that side stream's flag would never return (CUDA programming guide, lazy loading,
// "concurrent execution").  Everything that is issued between a chain launch and the end of its side-stream items is
// therefore loaded (and given its attributes) before the first chain launch on a device.
int chain_preload() {
    static std::mutex mu;
    static std::vector<int> done;
    int dev = 0;
    MPQR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    for (int d : done)
        if (d == dev) return MPQR_OK;
    MPQR_TRY(su_attrs<16>());
    MPQR_TRY(su_attrs<32>());
    (void)max_cluster();
    const void* fns[] = {(const void*)inpanel_s_kernel<16>, (const void*)inpanel_u_kernel<16>, (const void*)inpanel_s4_kernel<16>,
                         (const void*)inpanel_u4_kernel<16>, (const void*)chain_gate_kernel,
                         (const void*)panel_finalize_kernel,
                         // everything else a factorisation may launch for the first time while a gate kernel spins
                         (const void*)inpanel_s_kernel<32>, (const void*)inpanel_u_kernel<32>, (const void*)inpanel_s4_kernel<32>,
                         (const void*)inpanel_u4_kernel<32>, (const void*)tinv_kernel,
                         (const void*)panel_chain_kernel<1>, (const void*)panel_chain_kernel<2>, (const void*)panel_chain_kernel<4>,
                         (const void*)panel_chain_kernel<8>,
                         (const void*)panel_block_kernel<16, 1>, (const void*)panel_block_kernel<16, 2>, (const void*)panel_block_kernel<16, 4>,
                         (const void*)panel_block_kernel<16, 8>, (const void*)panel_block_kernel<32, 1>, (const void*)panel_block_kernel<32, 2>,
                         (const void*)panel_block_kernel<32, 4>};
    MPQR_TRY(func_attr_once((const void*)tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(((size_t)RMAX * TLD + 64 * 65 + (size_t)RMAX * RMAX) * sizeof(float))));
    {
        const void* ch[4] = {(const void*)panel_chain_kernel<1>, (const void*)panel_chain_kernel<2>, (const void*)panel_chain_kernel<4>,
                             (const void*)panel_chain_kernel<8>};
        const int rp[4] = {1, 2, 4, 8};
        for (int i = 0; i < 4; ++i) {
            MPQR_TRY(func_attr_once(ch[i], cudaFuncAttributeMaxDynamicSharedMemorySize, rp[i] * NT * 64));
            if (max_cluster() > 8) MPQR_TRY(func_attr_once(ch[i], cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
    }
    MPQR_TRY(preload_tc_gemm());
    MPQR_TRY(preload_simt_gemm());
    MPQR_TRY(preload_util_kernels());
    for (const void* f : fns) {
        cudaFuncAttributes fa;
        MPQR_CUDA(cudaFuncGetAttributes(&fa, f));
    }
    memop_api();
    done.push_back(dev);
    return MPQR_OK;
}

template <int RPT>
int launch_chain_t(const ChainArgs& a, int CS, cudaStream_t stream) {
    const size_t smem = (size_t)RPT * NT * 64;
    MPQR_TRY(func_attr_once((const void*)panel_chain_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) MPQR_TRY(func_attr_once((const void*)panel_chain_kernel<RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CS > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CS;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    at[na++] = pdl_attr();
    cfg.attrs = at;
    cfg.numAttrs = na;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, panel_chain_kernel<RPT>, a, CS));
    return MPQR_OK;
}
int launch_chain(const ChainArgs& a, int RPT, int CS, cudaStream_t st) {
    if (RPT == 1) return launch_chain_t<1>(a, CS, st);
    if (RPT == 2) return launch_chain_t<2>(a, CS, st);
    if (RPT == 4) return launch_chain_t<4>(a, CS, st);
    return launch_chain_t<8>(a, CS, st);
}

}  // namespace

bool panel_chain_ok(const PanelArgs& a) {
    if (getenv("MPQR_NO_CHAIN") || !a.chain_side || !a.chain_flags || !a.chain_ctr || !a.ws) return false;
    const int D = a.m - a.lam, pw = a.pw;
    if (pw < 32 || (pw & 15) || D < pw + 32 || (long)D > block_capacity(16) || a.ws_rows < D) return false;
    if ((a.lda & 3) || (reinterpret_cast<uintptr_t>(a.A + (size_t)a.lam * a.lda + a.acol) & 15)) return false;
    if (a.Y32 && ((a.ld32 & 3) || (reinterpret_cast<uintptr_t>(a.Y32) & 15) || a.lam != a.blk_row0)) return false;
    if (a.Y16 && ((a.ldy16 & 3) || (reinterpret_cast<uintptr_t>(a.Y16) & 7))) return false;
    if (a.force_b || a.dbg) return false;
    return true;
}

size_t panel_ws_bytes(long max_rows) {
    return ((size_t)max_rows * (2 * RMAX + 32) + 2 * 8 * 256 + (size_t)(NREP + 1) * RMAX * SLD + 4 + 256 + 3 * (size_t)RMAX * RMAX) * sizeof(float) +
           2 * (size_t)RMAX * RMAX * 2 + 256;
}

// PanelArgs output pointers address (row blk_row0, first panel column); zr = lam - blk_row0 rows
// above the panel are structurally zero.
int launch_panel(const PanelArgs& a, cudaStream_t stream, long* launches) {
    if (a.pw < 1 || a.pw > kPanelMaxWidth || a.lam < 0 || a.acol < 0 || a.lam >= a.m || a.blk_row0 > a.lam) {
        set_error("panel: bad arguments lam=%d pw=%d m=%d n=%d", a.lam, a.pw, a.m, a.n);
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    const int D = a.m - a.lam, pw = a.pw;
    // register-block width: 32 columns when two rows per thread are enough, else 16
    int B = (a.force_b == 16 || a.force_b == 32) ? a.force_b : ((pw > 16 && (long)D <= block_capacity(32)) ? 32 : 16);
    const bool chain = panel_chain_ok(a);  // persistent cluster kernel: 16-column register blocks for every D
    if (chain) B = 16;
    if ((long)D > block_capacity(B)) {
        if ((long)D <= block_capacity(16)) B = 16;
        else return launch_panel_legacy(a, stream, launches);
    }
    const int nblk = ceil_div(pw, B);
    const int zr = a.lam - a.blk_row0;
    float* Ablk = a.A + (size_t)a.lam * a.lda + a.acol;
    // outputs at panel row 0 (= global row lam)
    float* Y32l = a.Y32 ? a.Y32 + (size_t)zr * a.ld32 : nullptr;
    float* W32l = a.W32 ? a.W32 + (size_t)zr * a.ld32 : nullptr;
    char* Y16l = a.Y16 ? (char*)a.Y16 + (size_t)zr * a.ldy16 * 2 : nullptr;
    char* W16l = a.W16 ? (char*)a.W16 + (size_t)zr * a.ldw16 * 2 : nullptr;
    int rpt = 1, cs = 1;

    if (nblk == 1) {
        if (!pick_shape(B, D, a.force_cs, a.force_rpt, &rpt, &cs)) { set_error("panel: sizing error D=%d", D); return MPQR_EINVAL; }
        BlockArgs b{};
        b.A = Ablk; b.lda = a.lda; b.D = D; b.bw = pw;
        b.Y32 = {Y32l, a.ld32, zr}; b.W32 = {W32l, a.ld32, zr};
        b.Y16 = {Y16l, a.ldy16, zr}; b.W16 = {W16l, a.ldw16, zr};
        b.bf16 = a.bf16; b.T = a.T; b.ldt = a.ldt; b.dbg = a.dbg;
        if (a.prof) a.prof->begin(a.prof->ctx, 4, stream, 4.0 * D * pw * pw, 20.0 * D * pw);
        MPQR_TRY(launch_block(B, b, rpt, cs, stream));
        if (a.prof) a.prof->end(a.prof->ctx, stream);
        if (a.dbg_caps) { a.dbg_caps[0] = max_cluster(); a.dbg_caps[1] = cs; a.dbg_caps[2] = rpt; }
        if (launches) *launches += 1;
        return MPQR_OK;
    }

    if (!a.ws || a.ws_rows < D) {
        set_error("panel: workspace missing/too small (pw=%d needs %d blocks, D=%d, ws_rows=%ld)", pw, nblk, D, a.ws ? a.ws_rows : 0L);
        return MPQR_EINVAL;
    }
    const bool mixed = a.Y16 && a.W16;
    const bool need_t = a.T || a.W32 || a.W16;
    if (need_t && !mixed && a.W32 && !a.Y32) { set_error("panel: W32 needs Y32 on the FP32 path"); return MPQR_EINVAL; }
    if (mixed && !a.W32) { set_error("panel: the mixed path needs the FP32 W master"); return MPQR_EINVAL; }
    Ws w = carve(a.ws, a.ws_rows);
    // FP32 Y of the whole panel: caller's array if given, else one of the two workspace buffers (consecutive chain panels
    // alternate: panel p's finalize / side updates may still read its Y while panel p+1's kernel writes the other one)
    float* Yp = Y32l ? Y32l : ((chain && (a.chain_buf & 1)) ? w.Y32p2 : w.Y32p);
    const long ldyp = Y32l ? a.ld32 : RMAX;
    // deferred outputs (panel_finalize_kernel): full register blocks only, vector-aligned A and 16-bit Y
    const bool defer = !a.dbg && (pw % B) == 0 && D >= pw + 32 && ((a.lda & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(Ablk) & 15) == 0) && ((ldyp & 3) == 0) && ((reinterpret_cast<uintptr_t>(Yp) & 15) == 0) &&
                       (!Y16l || (((a.ldy16 & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y16) & 7) == 0)));
    auto finalize = [&](cudaStream_t fs) -> int {
        const int grid = sm_count(di) * 4;
        cudaLaunchAttribute pat[1] = {pdl_attr()};
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = fs; cfg.attrs = pat; cfg.numAttrs = 1;
        MPQR_CUDA(cudaLaunchKernelEx(&cfg, panel_finalize_kernel, (const float*)Yp, ldyp, Ablk, a.lda, (void*)a.Y16, a.ldy16, D, pw, B, zr, a.bf16));
        if (launches) *launches += 1;
        return MPQR_OK;
    };
    cudaStream_t ts = stream;   // everything after the register blocks (finalize, Gram / T / W)
    if (chain) {
        HostProfScope hp(0);
        if (!pick_shape(16, D, a.force_cs, a.force_rpt, &rpt, &cs)) { set_error("panel: sizing error D=%d", D); return MPQR_EINVAL; }
        const int nblocks = pw / 16;
        const int nfarb = nblocks - 2;  // blocks with a side update (the cluster applies block jb to block jb+1 itself)
        float* Tsl = w.Tsl + (size_t)(a.chain_buf & 1) * 8 * 256;
        ChainArgs ca{};
        ca.A = Ablk; ca.lda = a.lda; ca.D = D; ca.pw = pw;
        ca.Yp = Yp; ca.ldyp = ldyp;
        ca.Tslots = Tsl; ca.tstride = 256;
        ca.base = *a.chain_ctr;
        *a.chain_ctr += (unsigned)nblocks;
        ca.flag_done = a.chain_flags;
        ca.flag_far = a.chain_flags + 1;
        ca.flag_started = a.chain_flags + 3;
        // this panel's columns were last written by the previous panel's side updates (if it covered them): block 0 waits
        // for the last value posted so far
        ca.have_wait0 = (a.chain_last_far && *a.chain_last_far != 0) ? 1 : 0;
        ca.wait0 = ca.have_wait0 ? *a.chain_last_far : 0u;
        float* SrepA = w.Srep + (size_t)2 * RMAX * SLD;
        ca.srep[0] = SrepA; ca.srep[1] = SrepA + (size_t)2 * RMAX * SLD;
        ca.dbg = a.chain_dbg;
        MPQR_TRY(chain_preload());
        if (a.prof) a.prof->begin(a.prof->ctx, 4, stream, 4.0 * D * pw * 16 + 4.0 * D * 16 * 16 * (nblocks - 1), 8.0 * D * pw);
        MPQR_TRY(launch_chain(ca, rpt, cs, stream));  // issued BEFORE the side stream's waits (a wait never queues ahead of its producer)
        if (a.prof) a.prof->end(a.prof->ctx, stream);
        if (launches) *launches += 1;
        int side_sms = sm_count(di) - cs;  // the cluster keeps its SMs for the whole panel
        if (side_sms < 8) side_sms = 8;
        // Ordering of the side updates.  S(jb) is held back at STREAM level until the cluster has posted block jb
        // (cuStreamWaitValue32; a one-thread gate kernel where the driver lacks it); the side flag is posted by the U kernel's last
        // CTA (ticket counter): a stream write behind the kernel would add the kernel's completion and the memory operation's
        // own latency to every block of the chain, and nothing spins for a kernel-side post.  Stream memory operations cost
        // the issuing thread ~26 us each on B200, so the six writes per panel that this saves are 40 ms of host time at 32768^2.
        // Dropped after measurement: S kernels that wait for the cluster's flag themselves (issued ahead, CTAs spinning).  It
        // was slower (117.4 against 113.1 ms: waiting CTAs hold SMs the rest-stream GEMMs need) and [B200, r2i] deadlocked when
        // the first S kernel of a panel filled the partition before the 16-CTA cluster had been placed (16 free SMs in ONE GPC).
        for (int jb = 0; jb < nfarb; ++jb) {
            // block jb's reflectors -> the rest of the panel beyond block jb+1
            const int j0 = jb * 16, Dj = D - j0;
            const int cfirst = (j0 + 32 < pw) ? j0 + 32 : pw;   // first column (panel-relative)
            const int nfar = pw - cfirst;
            MPQR_TRY(stream_wait_geq(a.chain_side, a.chain_flags, ca.base + (unsigned)(jb + 1)));
            MPQR_TRY(launch_su<16>(Tsl + (size_t)jb * 256, Yp + (size_t)j0 * ldyp + j0, ldyp, Ablk + (size_t)j0 * a.lda + cfirst, a.lda, Dj,
                                   nfar, ca.srep[jb & 1], w.Sfin, side_sms, a.chain_side, launches, a.prof, false,
                                   a.chain_flags + 1, ca.base + (unsigned)(jb + 1), a.chain_flags + 2));
            if (a.chain_last_far) *a.chain_last_far = ca.base + (unsigned)(jb + 1);
        }
        if (a.prof) a.prof->begin(a.prof->ctx, 4, ts, 0.0, 0.0);
        MPQR_TRY(finalize(ts));
        if (a.prof) a.prof->end(a.prof->ctx, ts);
    }
    for (int j0 = 0; !chain && j0 < pw;) {
        const int Dj = D - j0;
        if (Dj <= 0) break;
        const int bw = (j0 + B < pw) ? B : pw - j0;
        if (!pick_shape(B, Dj, a.force_cs, a.force_rpt, &rpt, &cs)) { set_error("panel: sizing error D=%d", Dj); return MPQR_EINVAL; }
        const int nrest = pw - (j0 + bw);
        BlockArgs b{};
        b.A = Ablk + (size_t)j0 * a.lda + j0; b.lda = a.lda; b.D = Dj; b.bw = bw;
        b.Y32 = {Yp + (size_t)j0 * ldyp + j0, ldyp, j0 + (Y32l ? zr : 0)};
        if (nrest > 0) { b.T = w.Wj; b.ldt = B; }  // block T for the in-panel update
        if (Y16l) b.Y16 = {Y16l + ((size_t)j0 * a.ldy16 + j0) * 2, a.ldy16, j0 + zr};
        b.bf16 = a.bf16; b.dbg = a.dbg; b.defer_out = defer ? 1 : 0;
        if (nrest > 0) { b.zero_buf = w.Srep; b.zero_n = 2 * RMAX * SLD; }  // the following S kernel accumulates into two replicas
        if (a.prof) a.prof->begin(a.prof->ctx, 4, stream, 4.0 * Dj * bw * bw, 14.0 * Dj * bw);
        MPQR_TRY(launch_block(B, b, rpt, cs, stream));
        if (a.prof) a.prof->end(a.prof->ctx, stream);
        if (launches) *launches += 1;
        if (nrest > 0) {
            float* Arest = b.A + bw;
            if (B == 32) MPQR_TRY(launch_su<32>(w.Wj, b.Y32.p, ldyp, Arest, a.lda, Dj, nrest, w.Srep, w.Sfin, sm_count(di), stream, launches, a.prof));
            else MPQR_TRY(launch_su<16>(w.Wj, b.Y32.p, ldyp, Arest, a.lda, Dj, nrest, w.Srep, w.Sfin, sm_count(di), stream, launches, a.prof));
        }
        j0 += bw;
    }
    if (defer && !chain) {
        if (a.prof) a.prof->begin(a.prof->ctx, 4, stream, 0.0, 10.0 * D * pw);  // reads the FP32 Y, writes A and the 16-bit Y
        MPQR_TRY(finalize(stream));
        if (a.prof) a.prof->end(a.prof->ctx, stream);
    }
    if (a.dbg_caps) { a.dbg_caps[0] = max_cluster(); a.dbg_caps[1] = cs; a.dbg_caps[2] = rpt; }
    if (!need_t) return MPQR_OK;

    HostProfScope hp1(1);
    const size_t tsmem_max = ((size_t)RMAX * TLD + 64 * 65 + (size_t)RMAX * RMAX) * sizeof(float);
    MPQR_TRY(func_attr_once((const void*)tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem_max));
    // merged Gram / next-panel product: the columns right of the panel in the Y16 array must be the operand shadow
    const int gsn = (mixed && a.gs_ncols > 0 && a.gs_ncols <= RMAX && a.gs_S16 && a.T) ? a.gs_ncols : 0;
    const size_t tsmem = ((size_t)RMAX * TLD + 64 * 65 + (gsn ? (size_t)RMAX * RMAX : 0)) * sizeof(float);
    void* T16p = (char*)w.T16 + (size_t)(a.chain_buf & 1) * RMAX * RMAX * 2;
    const int Dz = D + zr;  // W is produced from the enclosing block's first row on (zero rows of Y give zero rows of W)
    float* Tdst = a.T ? a.T : w.T32;
    const int ldt = a.T ? a.ldt : RMAX;
    if (a.prof) a.prof->begin(a.prof->ctx, 6, ts, 4.0 * D * pw * pw, 10.0 * D * pw);
    if (mixed) {
        // Gram and W on tensor cores, from the 16-bit Y the trailing update uses
        const long ldg = gsn ? 2 * RMAX : RMAX;
        MPQR_TRY(tc_gemm_tn(Y16l, a.ldy16, Y16l, a.ldy16, w.G, ldg, pw, pw + gsn, D, a.bf16, 1, ts, launches));
        MPQR_TRY(launch_tinv(w.G, ldg, pw, Tdst, ldt, T16p, RMAX, a.bf16, tsmem, ts, gsn ? w.G + pw : nullptr, ldg, gsn, a.gs_S16, a.gs_lds16));
        if (launches) *launches += 1;
        if (!a.defer_w) MPQR_TRY(tc_gemm_nn_store(a.Y16, a.ldy16, T16p, RMAX, a.W32, a.ld32, a.W16, a.ldw16, Dz, pw, pw, a.bf16, ts, launches));
    } else {
        MPQR_TRY(sgemm_tn(Yp, ldyp, Yp, ldyp, w.G, RMAX, pw, pw, D, ts, launches));
        MPQR_TRY(launch_tinv(w.G, RMAX, pw, Tdst, ldt, nullptr, 0, 0, tsmem, ts));   // (FP32 path: plain Gram, W formed here)
        if (launches) *launches += 1;
        if (a.W32) {
            MPQR_TRY(sgemm_nn_store(a.Y32, a.ld32, Tdst, ldt, a.W32, a.ld32, Dz, pw, pw, ts));
            if (launches) *launches += 1;
        }
    }
    if (a.prof) a.prof->end(a.prof->ctx, ts);
    return MPQR_OK;
}


int chain_preload_all() { return chain_preload(); }

int panel_form_w(const PanelArgs& a, cudaStream_t stream, long* launches) {
    if (!a.Y16 || !a.W16 || !a.W32 || !a.ws) { set_error("panel_form_w: mixed-path arguments missing"); return MPQR_EINVAL; }
    HostProfScope hp1(1);
    Ws w = carve(a.ws, a.ws_rows);
    void* T16p = (char*)w.T16 + (size_t)(a.chain_buf & 1) * RMAX * RMAX * 2;
    const int Dz = a.m - a.blk_row0;
    return tc_gemm_nn_store(a.Y16, a.ldy16, T16p, RMAX, a.W32, a.ld32, a.W16, a.ldw16, Dz, a.pw, a.pw, a.bf16, stream, launches);
}

}  // namespace mpqr
