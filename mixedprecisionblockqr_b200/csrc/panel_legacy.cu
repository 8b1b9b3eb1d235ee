// panel_legacy.cu — multi-CTA shared-memory Householder panel kernel (sm_100a).  FALLBACK ONLY: used
// when a panel is taller than one thread-block cluster can hold in registers (D > 32768, see
// panel.cu); one grid-wide exchange through L2 per column makes it latency-bound.
//
// Replaces the reference's HOST panel factorisation h_householder_qr (Cuda/qr.cu:198-293,
// single CPU thread, forces a full-matrix PCIe round trip per panel, :1080-1082/:1215) and
// the 3r+2 launches of dev_wy_transform (Cuda/qr.cu:535-600, K1-K4 of SURVEY 2.4) by ONE
// persistent kernel per panel:
//
//   * the D x pw panel is distributed by rows over NC clusters x CS CTAs and stays resident in
//     shared memory for all pw reflector steps: HBM sees one coalesced read and one coalesced
//     write of the panel (algorithmic bytes 8*D*pw);
//   * per column ONE reduction: the dots g_j = u^T a_j (j >= k) give both the column norm (g_k)
//     and v^T a_j = g_j + s*mu*a_kj, so the rank-1 update of step k and the dots of step k+1
//     are fused into a single pass over the slice (4 rows in flight per warp for ILP);
//   * the reduction is hierarchical: inside a thread-block cluster the partial vectors are
//     all-gathered through distributed shared memory and one hardware cluster barrier; only
//     the cluster leaders exchange through L2 (per-leader slots + monotonic counter) and
//     broadcast the result back through DSMEM.  Panels that fit one cluster (all of C2/C3)
//     never touch L2 inside the column loop;
//   * tail (still in shared memory): Gram matrix Y^T Y -> T by the larft recurrence
//     (T[0:c,c] = -2 T[0:c,0:c] G[0:c,c], T[c,c] = 2) -> W = Y T, emitted as FP32 and as
//     the FP16/BF16 operands of the tensor-core trailing update.
//
// Conventions mirrored from the reference: sign = (u0 >= 0) ? +1 : -1
// (:229-235); zero column => reflector skipped (:242-244); unit vector w (beta = 2) stored
// one row below the diagonal (:283-285); R_kk = -sign*||u||.
#include "common.cuh"

namespace mpqr {
namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int RI = 4;                                       // rows in flight per warp
constexpr int CSMAX = 16;                                   // max cluster size
constexpr int WS_LD = kPanelMaxWidth;
constexpr int WS_ARRAY = kPanelMaxWidth * kPanelMaxWidth;   // floats of the Gram accumulator
constexpr int MAXNC = 160;                                  // max clusters of one launch (CS = 1: every CTA)
constexpr int SLOT_FLOATS = 2 * WS_LD;                      // per-cluster slot: partial dots | pivot row
// sync workspace (floats): slots[2][MAXNC][SLOT_FLOATS] | gram[WS_ARRAY] | counter
constexpr size_t WS_SLOTS = (size_t)2 * MAXNC * SLOT_FLOATS;

// ------------------------------------------------------------------ cluster / DSMEM PTX
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t remote_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote_addr), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Optional phase profiling (PanelArgs.dbg != null): CTA 0 / thread 0 accumulates clock64 deltas.
// dbg[0]=pass dbg[1]=local reduce+publish dbg[2]=cluster barrier(s)+leader exchange dbg[3]=gather
// dbg[4]=scalars dbg[5]=load dbg[6]=store dbg[7]=steps dbg[8]=G dbg[9]=rows_per_cta dbg[10]=gram+T
// dbg[11]=CS dbg[12]=NC
#define PROF_MARK(slot)                                               \
    if (prof) {                                                       \
        long long t__ = clock64();                                    \
        pacc[slot] += t__ - tprev;                                    \
        tprev = t__;                                                  \
    }

template <int CPL>
struct RowVec;
template <>
struct RowVec<1> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[1]) { x[0] = p[0]; }
    static __device__ __forceinline__ void store(float* p, const float (&x)[1]) { p[0] = x[0]; }
};
template <>
struct RowVec<2> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[2]) {
        float2 v = *reinterpret_cast<const float2*>(p);
        x[0] = v.x; x[1] = v.y;
    }
    static __device__ __forceinline__ void store(float* p, const float (&x)[2]) {
        *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]);
    }
};
template <>
struct RowVec<4> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[4]) {
        float4 v = *reinterpret_cast<const float4*>(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&x)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    }
};

template <int CPL>
__device__ __forceinline__ float pick(const float (&x)[CPL], int c) {
    float v = x[0];
#pragma unroll
    for (int q = 1; q < CPL; ++q) v = (c == q) ? x[q] : v;
    return v;
}

__device__ __forceinline__ void store16(void* base, long idx, float v, int bf16) {
    if (bf16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// Shared-memory layout (floats):
//   red[NW*PWP] | psum[PWP] | gsum[PWP] | prow[PWP] | diag[PWP] | gcol[2*PWP] | gt[PWP*(PWP+1)]
//   | xslot[2][CSMAX][PWP] | xprow[2][PWP] | fin[2][2*PWP] | pad to 4 | slice[rows*PWP]
template <int CPL>
__host__ __device__ constexpr int fixed_floats() {
    constexpr int PWP = 32 * CPL;
    int f = NW * PWP + 6 * PWP + PWP * (PWP + 1) + 2 * CSMAX * PWP + 2 * PWP + 4 * PWP;
    return (f + 3) & ~3;
}

template <int CPL>
__global__ void __launch_bounds__(NT, 1)
panel_kernel(PanelArgs a, int rows_per_cta, int use_smem, int CS, int NC) {
    constexpr int PWP = 32 * CPL;
    constexpr int GLD = PWP + 1;
    constexpr int NG = NT / PWP;                       // gather groups
    extern __shared__ __align__(16) float smem[];
    float* red = smem;
    float* psum = red + NW * PWP;
    float* gsum = psum + PWP;
    float* prow = gsum + PWP;
    float* diag = prow + PWP;
    float* gcol = diag + PWP;                // 2 * PWP (double buffered)
    float* gt = gcol + 2 * PWP;
    float* xslot = gt + PWP * GLD;           // [2][CSMAX][PWP]  written by cluster peers
    float* xprow = xslot + 2 * CSMAX * PWP;  // [2][PWP]         pivot row, written by its owner
    float* fin = xprow + 2 * PWP;            // [2][2*PWP]       final dots | pivot row from the leader
    float* slice_sm = smem + fixed_floats<CPL>();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = CS * NC;
    const unsigned crank = (CS > 1) ? cluster_ctarank() : 0u;
    const int cid = blockIdx.x / CS;  // cluster index
    const int lam = a.lam, pw = a.pw;
    const int D = a.m - lam;
    const int kr = pw < D ? pw : D;  // reflectors in this panel
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = (r0 + rows_per_cta < D) ? r0 + rows_per_cta : D;
    const int nrows = r1 > r0 ? r1 - r0 : 0;
    float* slice = use_smem ? slice_sm : a.scratch + (size_t)r0 * PWP;
    float* slots = a.sync_ws;
    float* gram_g = slots + WS_SLOTS;
    unsigned* ctr = reinterpret_cast<unsigned*>(gram_g + WS_ARRAY);
    const long lda = a.lda;
    float* Ablk = a.A + (size_t)lam * lda + a.acol;  // element (row lam, panel column 0)
    unsigned bar_id = 0;

    const bool prof = (a.dbg != nullptr) && blockIdx.x == 0 && tid == 0;
    long long pacc[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // registers: no memory traffic while timing
    long long tprev = prof ? clock64() : 0;
    // ---- load the slice (coalesced along the panel row), zero-pad columns >= pw
    for (int idx = tid; idx < nrows * PWP; idx += NT) {
        int li = idx / PWP, c = idx - li * PWP;
        slice[idx] = (c < pw) ? Ablk[(size_t)(r0 + li) * lda + c] : 0.f;
    }
    // the Gram accumulator is used (atomically) only after >= 1 grid-wide sync: zero it here
    if (G > 1)
        for (int idx = blockIdx.x * NT + tid; idx < WS_ARRAY; idx += G * NT) gram_g[idx] = 0.f;
    // peers must not write into our xslot before we are running: cluster-wide start barrier
    if (CS > 1) {
        cluster_arrive();
        cluster_wait();
    } else {
        __syncthreads();
    }
    PROF_MARK(5);

    float tau[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) tau[q] = 0.f;
    float smu = 0.f, vinv = 0.f;

    // step s: apply reflector s-1 (if s > 0) and accumulate the dots of column s (if s < kr)
    for (int step = 0; step <= kr; ++step) {
        const int kprev = step - 1;
        const bool do_upd = step > 0, do_dot = step < kr;
        const int lk = do_upd ? kprev / CPL : 0, ck = do_upd ? kprev % CPL : 0;
        const int ln = step / CPL, cn = step % CPL;
        const int par = step & 1;
        float acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = 0.f;

        int li0 = kprev - r0;  // first slice row touched by this step
        if (li0 < 0) li0 = 0;
        // fixed warp <-> row mapping (row li belongs to warp li % NW); RI rows in flight
        const int first = li0 + ((warp - li0) & (NW - 1));
        // columns < kprev (< step when there is no update) are final: their lanes stay out of
        // shared memory, which is the bandwidth that bounds this loop
        const bool lane_on = (lane * CPL + CPL - 1) >= (do_upd ? kprev : step);
        for (int lb = first; lb < nrows; lb += NW * RI) {
            float x[RI][CPL];
            bool ok[RI];
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                const int li = lb + u * NW;
                ok[u] = li < nrows;
                if (ok[u] && lane_on) {
                    RowVec<CPL>::load(slice + (size_t)li * PWP + lane * CPL, x[u]);
                } else {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) x[u][q] = 0.f;
                }
            }
            if (do_upd) {
                float xk[RI];
#pragma unroll
                for (int u = 0; u < RI; ++u) xk[u] = __shfl_sync(0xffffffffu, pick<CPL>(x[u], ck), lk);
#pragma unroll
                for (int u = 0; u < RI; ++u) {
                    const int i = r0 + lb + u * NW;
                    const float vi = (i == kprev) ? xk[u] + smu : xk[u];
#pragma unroll
                    for (int q = 0; q < CPL; ++q) x[u][q] = fmaf(-vi, tau[q], x[u][q]);
                    if (lane == lk) {
                        const float wv = vi * vinv;
#pragma unroll
                        for (int q = 0; q < CPL; ++q) x[u][q] = (q == ck) ? wv : x[u][q];
                    }
                    if (ok[u] && lane_on) RowVec<CPL>::store(slice + (size_t)(lb + u * NW) * PWP + lane * CPL, x[u]);
                }
            }
            if (do_dot) {
                float xn[RI];
#pragma unroll
                for (int u = 0; u < RI; ++u) xn[u] = __shfl_sync(0xffffffffu, pick<CPL>(x[u], cn), ln);
#pragma unroll
                for (int u = 0; u < RI; ++u) {
                    const int i = r0 + lb + u * NW;
                    const float xv = (ok[u] && i >= step) ? xn[u] : 0.f;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) acc[q] = fmaf(xv, x[u][q], acc[q]);
                }
            }
        }
        if (!do_dot) break;

#pragma unroll
        for (int q = 0; q < CPL; ++q) red[warp * PWP + lane * CPL + q] = acc[q];
        __syncthreads();
        PROF_MARK(0);
        if (tid < PWP) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[w * PWP + tid];
            if (G > 1) {
                psum[tid] = s;
            } else {
                gsum[tid] = s;
                prow[tid] = slice[(size_t)step * PWP + tid];
            }
        }
        if (G > 1) {
            __syncthreads();
            // ---- level 1: all-gather the partial vectors inside the cluster through DSMEM
            if (CS > 1) {
                for (int idx = tid; idx < CS * PWP; idx += NT) {
                    const int peer = idx / PWP, j = idx - peer * PWP;
                    if (j >= step)
                        st_cluster(map_to_cta(smem_addr(&xslot[(par * CSMAX + (int)crank) * PWP + j]), (unsigned)peer), psum[j]);
                }
                if (step >= r0 && step < r1) {  // this CTA owns the pivot row
                    for (int idx = tid; idx < CS * PWP; idx += NT) {
                        const int peer = idx / PWP, j = idx - peer * PWP;
                        st_cluster(map_to_cta(smem_addr(&xprow[par * PWP + j]), (unsigned)peer),
                                   slice[(size_t)(step - r0) * PWP + j]);
                    }
                }
                PROF_MARK(1);
                cluster_arrive();
                cluster_wait();
                if (tid < PWP) {
                    float t = 0.f;
                    if (tid >= step)
                        for (int c = 0; c < CS; ++c) t += xslot[(par * CSMAX + c) * PWP + tid];
                    psum[tid] = t;  // cluster sum (identical in every CTA of the cluster)
                }
                __syncthreads();
            } else {
                PROF_MARK(1);
            }
            if (NC == 1) {
                if (tid < PWP) {
                    gsum[tid] = psum[tid];
                    prow[tid] = xprow[par * PWP + tid];
                }
                PROF_MARK(2);
            } else {
                // ---- level 2: cluster leaders exchange through L2, then broadcast through DSMEM
                const int oc = (step / rows_per_cta) / CS;  // cluster that owns the pivot row
                ++bar_id;
                if (crank == 0) {
                    float* myslot = slots + ((size_t)par * MAXNC + cid) * SLOT_FLOATS;
                    if (tid < PWP) {
                        __stcg(&myslot[tid], psum[tid]);
                        if (cid == oc) {
                            const float pv = (CS > 1) ? xprow[par * PWP + tid] : slice[(size_t)(step - r0) * PWP + tid];
                            __stcg(&myslot[WS_LD + tid], pv);
                        }
                    }
                    __syncthreads();
                    if (tid == 0) {
                        asm volatile("fence.acq_rel.gpu;" ::: "memory");
                        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
                        const unsigned target = a.ctr_base + (unsigned)NC * bar_id;
                        while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
                        }
                    }
                    __syncthreads();
                    const float* sl = slots + (size_t)par * MAXNC * SLOT_FLOATS;
                    const int j = tid % PWP, grp = tid / PWP;
                    float s = 0.f;
                    if (j >= step) {
                        for (int c0 = grp; c0 < NC; c0 += NG * 8) {  // 8 L2 loads in flight
                            float v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int c = c0 + u * NG;
                                v[u] = (c < NC) ? __ldcg(&sl[(size_t)c * SLOT_FLOATS + j]) : 0.f;
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) s += v[u];
                        }
                    }
                    red[grp * PWP + j] = s;
                    float pv = 0.f;
                    if (tid < PWP) pv = __ldcg(&sl[(size_t)oc * SLOT_FLOATS + WS_LD + tid]);
                    __syncthreads();
                    if (tid < PWP) {
                        float t = 0.f;
#pragma unroll
                        for (int gq = 0; gq < NG; ++gq) t += red[gq * PWP + tid];
                        if (CS > 1) {
                            fin[par * 2 * PWP + tid] = t;
                            fin[par * 2 * PWP + PWP + tid] = pv;
                        } else {
                            gsum[tid] = t;
                            prow[tid] = pv;
                        }
                    }
                    if (CS > 1) {
                        __syncthreads();
                        for (int idx = tid; idx < (CS - 1) * 2 * PWP; idx += NT) {
                            const int peer = 1 + idx / (2 * PWP), j2 = idx % (2 * PWP);
                            st_cluster(map_to_cta(smem_addr(&fin[par * 2 * PWP + j2]), (unsigned)peer), fin[par * 2 * PWP + j2]);
                        }
                    }
                }
                if (CS > 1) {
                    cluster_arrive();
                    cluster_wait();
                    if (tid < PWP) {
                        gsum[tid] = fin[par * 2 * PWP + tid];
                        prow[tid] = fin[par * 2 * PWP + PWP + tid];
                    }
                }
                PROF_MARK(2);
            }
        } else {
            PROF_MARK(1);
        }
        __syncthreads();
        PROF_MARK(3);

        // reflector scalars (every thread, redundantly): MUFU.RSQ + one Newton step each, i.e.
        // full FP32 accuracy at a fraction of the latency of IEEE sqrt/div
        const float gk = gsum[step], ak = prow[step];
        const bool skip = !(gk > 0.f);
        const float rs = rsqrtf(skip ? 1.f : gk);
        float mu = gk * rs;
        mu = fmaf(0.5f * rs, fmaf(-mu, mu, gk), mu);  // sqrt(gk)
        if (skip) mu = 0.f;
        smu = (ak >= 0.f) ? mu : -mu;
        const float vn2 = 2.f * mu * (mu + fabsf(ak));
        float rv = rsqrtf(skip ? 1.f : vn2);
        rv = rv * fmaf(-0.5f * vn2, rv * rv, 1.5f);   // 1/sqrt(vn2)
        vinv = skip ? 0.f : rv;
        const float inv2 = skip ? 0.f : 2.f * rv * rv;  // 2/vn2
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            int col = lane * CPL + q;
            tau[q] = (col > step && col < pw) ? (gsum[col] + smu * prow[col]) * inv2 : 0.f;
        }
        if (skip) smu = 0.f;
        if (tid == 0) diag[step] = skip ? ak : -smu;
        PROF_MARK(4);
        // (gsum/prow are rewritten only after the next pass's __syncthreads)
    }
    __syncthreads();

    // ---- packed output: R above the diagonal, R_kk on it, w shifted one row down
    for (int idx = tid; idx < nrows * PWP; idx += NT) {
        int li = idx / PWP, c = idx - li * PWP;
        if (c >= pw) continue;
        int i = r0 + li;
        float v = slice[idx];
        if (i < c) {
            Ablk[(size_t)i * lda + c] = v;
        } else {
            Ablk[(size_t)(i + 1) * lda + c] = v;
            if (i == c) Ablk[(size_t)i * lda + c] = diag[c];
        }
    }
    __syncthreads();
    PROF_MARK(6);
    if (prof) {
        pacc[7] += kr; a.dbg[8] = G; a.dbg[9] = rows_per_cta; a.dbg[11] = CS; a.dbg[12] = NC;
    }

    const bool want16y = a.Y16 != nullptr, want16w = a.W16 != nullptr;
    const bool need_t = a.T || a.W32 || want16w;
    const int rofs = lam - a.blk_row0;  // output row of panel row 0
    if (a.Y32 || want16y || need_t) {
        // ---- slice := Y (zero strictly above the diagonal and for columns without reflector)
        for (int idx = tid; idx < nrows * PWP; idx += NT) {
            int li = idx / PWP, c = idx - li * PWP;
            int i = r0 + li;
            if (i < c || c >= kr) slice[idx] = 0.f;
        }
        __syncthreads();
        for (int idx = tid; idx < nrows * PWP; idx += NT) {
            int li = idx / PWP, c = idx - li * PWP;
            if (c >= pw) continue;
            long orow = rofs + r0 + li;
            float v = slice[idx];
            if (a.Y32) a.Y32[orow * a.ld32 + c] = v;
            if (want16y) store16(a.Y16, orow * a.ldy16 + c, v, a.bf16);
        }
        // rows [blk_row0, lam) of the outputs are structurally zero
        for (long idx = (long)blockIdx.x * NT + tid; idx < (long)rofs * pw; idx += (long)G * NT) {
            long rr = idx / pw;
            int c = (int)(idx - rr * pw);
            if (a.Y32) a.Y32[rr * a.ld32 + c] = 0.f;
            if (a.W32) a.W32[rr * a.ld32 + c] = 0.f;
            if (want16y) store16(a.Y16, rr * a.ldy16 + c, 0.f, a.bf16);
            if (want16w) store16(a.W16, rr * a.ldw16 + c, 0.f, a.bf16);
        }
    }
    if (need_t) {
        // ---- Gram matrix G[t][c] = sum_i y_it y_ic (strict upper part is what T needs)
        for (int idx = tid; idx < PWP * GLD; idx += NT) gt[idx] = 0.f;
        __syncthreads();
        {
            constexpr int NTC = PWP / 4, NTR = PWP / 8;
            if (tid < NTR * NTC) {
                const int tr = tid / NTC, tc = tid - tr * NTC;
                if (8 * tr < 4 * tc + 3) {  // tile contains at least one (t < c)
                    float g[8][4];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) g[u][v] = 0.f;
                    int lstart = 4 * tc - r0;  // y_ic = 0 for i < c
                    if (lstart < 0) lstart = 0;
                    for (int li = lstart; li < nrows; ++li) {
                        const float* row = slice + (size_t)li * PWP;
                        float4 t0 = *reinterpret_cast<const float4*>(row + 8 * tr);
                        float4 t1 = *reinterpret_cast<const float4*>(row + 8 * tr + 4);
                        float4 cc = *reinterpret_cast<const float4*>(row + 4 * tc);
                        float yt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                        float yc[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                        for (int u = 0; u < 8; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) g[u][v] = fmaf(yt[u], yc[v], g[u][v]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            int t = 8 * tr + u, c = 4 * tc + v;
                            if (t < c && c < kr) {
                                if (G > 1) atomicAdd(&gram_g[t * WS_LD + c], g[u][v]);
                                else gt[t * GLD + c] = g[u][v];
                            }
                        }
                }
            }
        }
        if (G > 1) {
            // grid-wide sync: cluster barrier, leaders through L2, cluster barrier
            if (CS > 1) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");  // order this CTA's global atomics
                cluster_arrive();
                cluster_wait();
            } else {
                __syncthreads();
            }
            if (NC > 1) {
                ++bar_id;
                if (crank == 0 && tid == 0) {
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
                    const unsigned target = a.ctr_base + (unsigned)NC * bar_id;
                    while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
                    }
                }
                if (CS > 1) {
                    cluster_arrive();
                    cluster_wait();
                } else {
                    __syncthreads();
                }
            }
            for (int idx = tid; idx < PWP * PWP; idx += NT) {
                int t = idx / PWP, c = idx - t * PWP;
                if (t < c && c < kr) gt[t * GLD + c] = __ldcg(&gram_g[t * WS_LD + c]);
            }
        }
        __syncthreads();

        // ---- T in place: column c of gt goes from G[0:c,c] to T[0:c,c]
        for (int c = 0; c < kr; ++c) {
            float* gc = gcol + (c & 1) * PWP;
            if (tid < c) gc[tid] = gt[tid * GLD + c];
            __syncthreads();
            const int t = tid >> 2, part = tid & 3;
            float s = 0.f;
            if (t < c)
                for (int u = t + part; u < c; u += 4) s = fmaf(gt[t * GLD + u], gc[u], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0 && t < c) gt[t * GLD + c] = -2.f * s;
            if (tid == c) gt[c * GLD + c] = 2.f;
        }
        __syncthreads();
        if (a.T && blockIdx.x == 0) {
            for (int idx = tid; idx < pw * pw; idx += NT) {
                int t = idx / pw, c = idx - t * pw;
                a.T[(size_t)t * a.ldt + c] = (t <= c && c < kr) ? gt[t * GLD + c] : 0.f;
            }
        }
        PROF_MARK(10);
        if (a.W32 || want16w) {
            // ---- W = Y T on the slice rows; lane <-> columns lane + 32 q (conflict-free T reads)
            for (int base = warp * 8; base < nrows; base += NW * 8) {
                float w[8][CPL];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr)
#pragma unroll
                    for (int q = 0; q < CPL; ++q) w[rr][q] = 0.f;
                int tmax = r0 + base + 8;  // y_it = 0 for t > i
                if (tmax > kr) tmax = kr;
                for (int t = 0; t < tmax; ++t) {
                    float tt[CPL];
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        int col = lane + 32 * q;
                        tt[q] = (col >= t) ? gt[t * GLD + col] : 0.f;
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        int li = base + rr;
                        float y = (li < nrows) ? slice[(size_t)li * PWP + t] : 0.f;
#pragma unroll
                        for (int q = 0; q < CPL; ++q) w[rr][q] = fmaf(y, tt[q], w[rr][q]);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    int li = base + rr;
                    if (li >= nrows) continue;
                    long orow = rofs + r0 + li;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        int col = lane + 32 * q;
                        if (col >= pw) continue;
                        if (a.W32) a.W32[orow * a.ld32 + col] = w[rr][q];
                        if (want16w) store16(a.W16, orow * a.ldw16 + col, w[rr][q], a.bf16);
                    }
                }
            }
        }
    }
    if (prof) {
        for (int i = 0; i < 11; ++i) a.dbg[i] += pacc[i];
    }
    // a CTA's shared memory must stay alive until no peer can write into it any more
    if (CS > 1) {
        cluster_arrive();
        cluster_wait();
    }
}

struct ClusterCaps {
    int max_cs;    // largest usable cluster size (16, 8 or 1)
    int max_nc16;  // co-resident clusters of 16 at full dynamic smem
    int max_nc8;   // co-resident clusters of 8
};

template <int CPL>
int query_caps(const DeviceInfo& di, ClusterCaps* out) {
    static ClusterCaps caps = {0, 0, 0};
    static bool done = false;
    if (!done) {
        MPQR_CUDA(cudaFuncSetAttribute(panel_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        cudaError_t e = cudaFuncSetAttribute(panel_kernel<CPL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        const bool np_ok = (e == cudaSuccess);
        if (!np_ok) cudaGetLastError();
        auto occ = [&](int cs) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(cs * 4);
            cfg.blockDim = dim3(NT);
            cfg.dynamicSmemBytes = di.max_smem_optin;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, panel_kernel<CPL>, &cfg) != cudaSuccess) {
                cudaGetLastError();
                n = 0;
            }
            return n;
        };
        caps.max_nc16 = np_ok ? occ(16) : 0;
        caps.max_nc8 = occ(8);
        caps.max_cs = caps.max_nc16 > 0 ? 16 : (caps.max_nc8 > 0 ? 8 : 1);
        done = true;
    }
    *out = caps;
    return MPQR_OK;
}

template <int CPL>
int launch_t(const PanelArgs& a, cudaStream_t stream, const DeviceInfo& di) {
    constexpr int PWP = 32 * CPL;
    const int D = a.m - a.lam;
    const size_t fixed = (size_t)fixed_floats<CPL>() * sizeof(float);
    const int max_rows_smem = (int)(((size_t)di.max_smem_optin - fixed - 256) / (PWP * sizeof(float)));
    const int cap = max_rows_smem - (max_rows_smem % NW);
    ClusterCaps caps;
    MPQR_TRY(query_caps<CPL>(di, &caps));
    if (a.force_cs > 0 && a.force_cs < caps.max_cs) caps.max_cs = a.force_cs;
    if (a.dbg_caps) { a.dbg_caps[0] = caps.max_cs; a.dbg_caps[1] = caps.max_nc16; a.dbg_caps[2] = caps.max_nc8; }
    int rows_per_cta, CS = 1, NC = 1, use_smem = 1;
    if (D <= cap && (D <= 256 || caps.max_cs == 1)) {
        rows_per_cta = D;
    } else {
        // ~96 rows per CTA (the pass costs ~10 cycles/row/step with 4 rows in flight, a cluster
        // barrier ~400), but never more than one cluster unless capacity forces it: a single
        // cluster never touches L2 inside the column loop
        const int need = ceil_div(D, cap);  // CTAs needed for capacity
        int want = ceil_div(D, a.rows_hint > 0 ? a.rows_hint : 96);
        if (want < need) want = need;
        if (need <= caps.max_cs && !(a.force_cs == 1)) {
            // fits one cluster: DSMEM all-gather + one hardware cluster barrier per column
            // (measured ~1.0k cycles vs 4-8k for any exchange through L2, profiles/r1_panel_probe.txt)
            CS = 1;
            while (CS < want && CS < caps.max_cs) CS *= 2;
            while (CS < need) CS *= 2;
        } else {
            // too tall for one cluster: flat exchange through L2 over as many CTAs as possible
            // (the pass is issue-bound, ~10 cycles per row per step, so rows per CTA must be small)
            CS = 1;
            NC = ceil_div(D, a.rows_hint > 0 ? a.rows_hint : 64);
            if (NC > sm_count(di)) NC = sm_count(di);  // cooperative grid: co-resident inside the issuing stream's SM partition
            if (NC > MAXNC) NC = MAXNC;
            if (NC < need) {
                use_smem = 0;
                if (!a.scratch || a.scratch_rows < D) {
                    set_error("panel: scratch buffer missing/too small for D=%d", D);
                    return MPQR_EINVAL;
                }
            }
        }
        rows_per_cta = round_up(ceil_div(D, CS * NC), NW);
        if (use_smem && rows_per_cta > cap) {
            set_error("panel: internal sizing error D=%d CS=%d NC=%d rows=%d cap=%d", D, CS, NC, rows_per_cta, cap);
            return MPQR_EINVAL;
        }
    }
    const int G = CS * NC;
    size_t smem = fixed + (use_smem ? (size_t)rows_per_cta * PWP * sizeof(float) : 0);
    PanelArgs args = a;
    if (NC > 1) {
        // L2 barriers executed by the leaders: one per reflector + one for the Gram reduction
        const int kr = a.pw < D ? a.pw : D;
        const bool need_t = a.T || a.W32 || a.W16;
        unsigned nbar = (unsigned)kr + (need_t ? 1u : 0u);
        if (!a.host_ctr) {
            set_error("panel: host_ctr missing");
            return MPQR_EINVAL;
        }
        args.ctr_base = *a.host_ctr;
        *a.host_ctr += (unsigned)NC * nbar;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(G);
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
    if (NC > 1) {
        // all clusters must be co-resident (leaders spin on an L2 counter)
        at[na].id = cudaLaunchAttributeCooperative;
        at[na].val.cooperative = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, panel_kernel<CPL>, args, rows_per_cta, use_smem, CS, NC));
    return MPQR_OK;
}

}  // namespace

size_t panel_sync_ws_bytes() { return (WS_SLOTS + WS_ARRAY) * sizeof(float) + 256; }
size_t panel_scratch_bytes(int max_rows) { return (size_t)max_rows * kPanelMaxWidth * sizeof(float); }

int launch_panel_legacy(const PanelArgs& a, cudaStream_t stream, long* launches) {
    if (a.pw < 1 || a.pw > kPanelMaxWidth || a.lam < 0 || a.acol < 0 || a.lam >= a.m ||
        a.blk_row0 > a.lam) {
        set_error("panel: bad arguments lam=%d pw=%d m=%d n=%d", a.lam, a.pw, a.m, a.n);
        return MPQR_EINVAL;
    }
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    int rc;
    if (a.pw <= 32) rc = launch_t<1>(a, stream, di);
    else if (a.pw <= 64) rc = launch_t<2>(a, stream, di);
    else rc = launch_t<4>(a, stream, di);
    if (rc == MPQR_OK && launches) *launches += 1;
    return rc;
}

}  // namespace mpqr
