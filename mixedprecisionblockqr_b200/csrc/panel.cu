// panel.cu — multi-CTA Householder panel factorisation + fused compact-WY (sm_100a).
//
// Replaces the reference's HOST panel factorisation h_householder_qr (Cuda/qr.cu:198-293,
// single CPU thread, forces a full-matrix PCIe round trip per panel, :1080-1082/:1215) and
// the 3r+2 launches of dev_wy_transform (Cuda/qr.cu:535-600, K1-K4 of SURVEY 2.4) by ONE
// persistent kernel per panel:
//
//   * the D x pw panel is distributed by rows over G CTAs and stays resident in shared
//     memory (<= ~150 KB/CTA) for all pw reflector steps: HBM sees one coalesced read and
//     one coalesced write of the panel (algorithmic bytes 8*D*pw);
//   * per column ONE grid-wide reduction: the dots g_j = u^T a_j (j >= k) give both the
//     column norm (g_k) and v^T a_j = g_j + s*mu*a_kj, so the rank-1 update of step k and
//     the dots of step k+1 are fused into a single pass over the slice;
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
constexpr int WS_LD = kPanelMaxWidth;                       // row stride of the global sync arrays
constexpr int WS_ARRAY = kPanelMaxWidth * kPanelMaxWidth;   // floats of the Gram accumulator
constexpr int MAXG = 160;                                   // max CTAs of one panel launch (>= #SMs)
constexpr int SLOT_FLOATS = 2 * WS_LD;                      // per-CTA slot: partial dots | pivot row
// sync workspace (floats): slots[2][MAXG][SLOT_FLOATS] | gram[WS_ARRAY] | counter
constexpr size_t WS_SLOTS = (size_t)2 * MAXG * SLOT_FLOATS;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (cooperatively launched, co-resident) grid arrive.  The counter only ever
// increases (across barriers AND across launches: the host passes the base), so it is never
// reset; comparisons are wrap-safe.  Release: bar.sync orders the CTA's slot stores before
// thread 0's fence + red; acquire: ld.acquire by thread 0, then bar.sync, then .cg loads.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
        }
    }
    __syncthreads();
}

// Optional phase profiling (PanelArgs.dbg != null): CTA 0 / thread 0 accumulates clock64 deltas.
// dbg[0]=pass dbg[1]=local reduce+publish dbg[2]=grid barrier dbg[3]=gather dbg[4]=scalars
// dbg[5]=load dbg[6]=store+tail dbg[7]=steps dbg[8]=G dbg[9]=rows_per_cta
#define PROF_MARK(slot)                                               \
    if (prof) {                                                       \
        long long t__ = clock64();                                    \
        a.dbg[slot] += t__ - tprev;                                   \
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

// Shared-memory layout (floats): red[NW*PWP] | gsum[PWP] | prow[PWP] | diag[PWP] | gcol[2*PWP]
//                                | gt[PWP*(PWP+1)] | pad to 4 | slice[rows*PWP]
template <int CPL>
__host__ __device__ constexpr int fixed_floats() {
    constexpr int PWP = 32 * CPL;
    int f = NW * PWP + 5 * PWP + PWP * (PWP + 1);
    return (f + 3) & ~3;
}

template <int CPL>
__global__ void __launch_bounds__(NT, 1)
panel_kernel(PanelArgs a, int rows_per_cta, int use_smem, int G) {
    constexpr int PWP = 32 * CPL;
    constexpr int GLD = PWP + 1;
    extern __shared__ __align__(16) float smem[];
    float* red = smem;
    float* gsum = red + NW * PWP;
    float* prow = gsum + PWP;
    float* diag = prow + PWP;
    float* gcol = diag + PWP;  // 2 * PWP (double buffered)
    float* gt = gcol + 2 * PWP;
    float* slice_sm = smem + fixed_floats<CPL>();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    constexpr int NG = NT / PWP;  // gather groups
    const long lda = a.lda;
    float* Ablk = a.A + (size_t)lam * lda + a.acol;  // element (row lam, panel column 0)

    const bool prof = (a.dbg != nullptr) && blockIdx.x == 0 && tid == 0;
    long long tprev = prof ? clock64() : 0;
    // ---- load the slice (coalesced along the panel row), zero-pad columns >= pw
    for (int idx = tid; idx < nrows * PWP; idx += NT) {
        int li = idx / PWP, c = idx - li * PWP;
        slice[idx] = (c < pw) ? Ablk[(size_t)(r0 + li) * lda + c] : 0.f;
    }
    // the Gram accumulator is used (atomically) only after >= 1 grid barrier: zero it here
    if (G > 1)
        for (int idx = blockIdx.x * NT + tid; idx < WS_ARRAY; idx += G * NT) gram_g[idx] = 0.f;
    __syncthreads();
    PROF_MARK(5);

    unsigned bar_id = 0;
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
        float acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = 0.f;

        int li0 = kprev - r0;  // first slice row touched by this step
        if (li0 < 0) li0 = 0;
        // keep the warp <-> row mapping fixed: start at the first row >= li0 owned by this warp
        int first = li0 + ((warp - li0) % NW + NW) % NW;
        for (int li = first; li < nrows; li += NW) {
            const int i = r0 + li;
            float x[CPL];
            float* p = slice + (size_t)li * PWP + lane * CPL;
            RowVec<CPL>::load(p, x);
            if (do_upd) {
                float xk = __shfl_sync(0xffffffffu, pick<CPL>(x, ck), lk);
                float vi = (i == kprev) ? xk + smu : xk;
#pragma unroll
                for (int q = 0; q < CPL; ++q) x[q] = fmaf(-vi, tau[q], x[q]);
                if (lane == lk) {
                    float wv = vi * vinv;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) x[q] = (q == ck) ? wv : x[q];
                }
                RowVec<CPL>::store(p, x);
            }
            if (do_dot && i >= step) {
                float xn = __shfl_sync(0xffffffffu, pick<CPL>(x, cn), ln);
#pragma unroll
                for (int q = 0; q < CPL; ++q) acc[q] = fmaf(xn, x[q], acc[q]);
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
                // publish this CTA's partial dots (and the pivot row if it owns it) in its slot;
                // slots are double-buffered by step parity and fully rewritten, never reset
                float* myslot = slots + ((size_t)(step & 1) * MAXG + blockIdx.x) * SLOT_FLOATS;
                __stcg(&myslot[tid], s);
                if (step >= r0 && step < r1) __stcg(&myslot[WS_LD + tid], slice[(size_t)(step - r0) * PWP + tid]);
            } else {
                gsum[tid] = s;
                prow[tid] = slice[(size_t)step * PWP + tid];
            }
        }
        PROF_MARK(1);
        if (G > 1) {
            grid_barrier(ctr, a.ctr_base + (unsigned)G * (++bar_id));
            PROF_MARK(2);
            // deterministic gather: group grp sums CTAs grp, grp+NG, ... for column j.
            // Loads are issued in batches of 8 before the first use (each is an L2 round trip).
            const int j = tid % PWP, grp = tid / PWP;
            const float* sl = slots + (size_t)(step & 1) * MAXG * SLOT_FLOATS;
            float s = 0.f;
            if (j >= step) {
                for (int c0 = grp; c0 < G; c0 += NG * 8) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        int c = c0 + u * NG;
                        v[u] = (c < G) ? __ldcg(&sl[(size_t)c * SLOT_FLOATS + j]) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) s += v[u];
                }
            }
            red[grp * PWP + j] = s;
            if (tid < PWP) prow[tid] = __ldcg(&sl[(size_t)(step / rows_per_cta) * SLOT_FLOATS + WS_LD + tid]);
            __syncthreads();
            if (tid < PWP) {
                float t = 0.f;
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) t += red[gq * PWP + tid];
                gsum[tid] = t;
            }
        }
        __syncthreads();
        PROF_MARK(3);

        // reflector scalars (every thread, redundantly)
        const float gk = gsum[step], ak = prow[step];
        const bool skip = !(gk > 0.f);
        const float mu = sqrtf(gk);
        smu = (ak >= 0.f) ? mu : -mu;
        const float vn2 = 2.f * mu * (mu + fabsf(ak));
        const float inv2 = skip ? 0.f : 2.f / vn2;
        vinv = skip ? 0.f : 1.f / sqrtf(vn2);
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
    if (prof) { a.dbg[7] += kr; a.dbg[8] = G; a.dbg[9] = rows_per_cta; }
    const bool want16y = a.Y16 != nullptr, want16w = a.W16 != nullptr;
    const bool need_t = a.T || a.W32 || want16w;
    if (!(a.Y32 || want16y || need_t)) return;

    // ---- slice := Y (zero strictly above the diagonal and for columns without reflector)
    for (int idx = tid; idx < nrows * PWP; idx += NT) {
        int li = idx / PWP, c = idx - li * PWP;
        int i = r0 + li;
        if (i < c || c >= kr) slice[idx] = 0.f;
    }
    __syncthreads();
    const int rofs = lam - a.blk_row0;  // output row of panel row 0
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
    if (!need_t) return;

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
                int lstart = 4 * tc - r0;  // y_ic = 0 for i < c  => rows below 4*tc contribute nothing
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
        grid_barrier(ctr, a.ctr_base + (unsigned)G * (++bar_id));
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
    if (!(a.W32 || want16w)) return;

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

template <int CPL>
int launch_t(const PanelArgs& a, cudaStream_t stream, const DeviceInfo& di) {
    constexpr int PWP = 32 * CPL;
    const int D = a.m - a.lam;
    const size_t fixed = (size_t)fixed_floats<CPL>() * sizeof(float);
    const int max_rows_smem = (int)(((size_t)di.max_smem_optin - fixed - 256) / (PWP * sizeof(float)));
    int rows_per_cta, G, use_smem = 1;
    if (D <= 512 && D <= max_rows_smem) {
        rows_per_cta = D;
        G = 1;
    } else {
        // measured (tools/panel_probe.py, profiles/r1_panel_probe.txt): the local pass costs
        // ~16 cycles/row/step while the gather grows only ~8 cycles per extra CTA -> use many CTAs
        rows_per_cta = ceil_div(D, di.num_sms);
        if (a.rows_hint > 0) rows_per_cta = a.rows_hint;
        if (rows_per_cta < 64) rows_per_cta = 64;
        rows_per_cta = round_up(rows_per_cta, NW);
        const int cap = max_rows_smem - (max_rows_smem % NW);
        if (rows_per_cta > cap) rows_per_cta = cap;
        if (ceil_div(D, rows_per_cta) > di.num_sms) rows_per_cta = round_up(ceil_div(D, di.num_sms), NW);
        if (rows_per_cta > max_rows_smem) {
            // does not fit: keep the slice in a global scratch buffer (L2-resident for moderate D)
            use_smem = 0;
            rows_per_cta = round_up(ceil_div(D, di.num_sms), NW);
            if (!a.scratch || a.scratch_rows < D) {
                set_error("panel: scratch buffer missing/too small for D=%d", D);
                return MPQR_EINVAL;
            }
        }
        G = ceil_div(D, rows_per_cta);
    }
    if (G > 1 && !di.coop) {
        set_error("panel: device lacks cooperative launch");
        return MPQR_ECUDA;
    }
    size_t smem = fixed + (use_smem ? (size_t)rows_per_cta * PWP * sizeof(float) : 0);
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        MPQR_CUDA(cudaFuncSetAttribute(panel_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       di.max_smem_optin));
        attr_set = true;
    }
    if (G > MAXG) {
        set_error("panel: %d CTAs exceed the sync workspace (MAXG=%d)", G, MAXG);
        return MPQR_EINVAL;
    }
    PanelArgs args = a;
    if (G > 1) {
        // grid barriers executed by this launch: one per reflector + one for the Gram reduction
        const int kr = a.pw < D ? a.pw : D;
        const bool need_t = a.T || a.W32 || a.W16;
        unsigned nbar = (unsigned)kr + (need_t ? 1u : 0u);
        if (!a.host_ctr) {
            set_error("panel: host_ctr missing");
            return MPQR_EINVAL;
        }
        args.ctr_base = *a.host_ctr;
        *a.host_ctr += (unsigned)G * nbar;
    }
    if (G > 1) {
        void* kargs[] = {(void*)&args, (void*)&rows_per_cta, (void*)&use_smem, (void*)&G};
        MPQR_CUDA(cudaLaunchCooperativeKernel((void*)panel_kernel<CPL>, dim3(G), dim3(NT), kargs, smem, stream));
    } else {
        panel_kernel<CPL><<<1, NT, smem, stream>>>(args, rows_per_cta, use_smem, G);
        MPQR_CUDA(cudaGetLastError());
    }
    return MPQR_OK;
}

}  // namespace

size_t panel_sync_ws_bytes() { return (WS_SLOTS + WS_ARRAY) * sizeof(float) + 256; }
size_t panel_scratch_bytes(int max_rows) { return (size_t)max_rows * kPanelMaxWidth * sizeof(float); }

int launch_panel(const PanelArgs& a, cudaStream_t stream, long* launches) {
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
