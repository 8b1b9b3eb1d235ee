// gemm_simt.cu — FP32 CUDA-core GEMMs of the FP32 driver (the "FP32-vs-FP32, tighter" path).
//
// Replace, for the FP32 sibling dev_block_qr_wy (reference Cuda/qr.cu:958-1047):
//   shared_mem_mmult_in_place_transpose_a + dev_cpy_strided_array (Cuda/mmult.cu:236-288,
//   Cuda/mmult.cuh:104-151) and dev_apply_qpanel_to_q (Cuda/qr.cu:843-855),
// but in factored form: S = W^T A22 (K = D rows) then A22 -= Y S (K = panel width), i.e.
// 4*D*N'*r flops per panel instead of the reference's 2*D^2*N' with a dense panel-Q.
#include "common.cuh"

namespace mpqr {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int GT = (BM / TM) * (BN / TN);  // 256 threads

// C tile = sum_k A_op[k][i] * B[k][j].
// TRANS_A = true : X stored [K x M] (row k contiguous in i)  -> S = X^T Z
// TRANS_A = false: X stored [M x K]                            -> C -= X S
// MODE 0: store, 1: subtract from C, 2: atomicAdd (split-K over blockIdx.z)
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st16(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void st16(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st16(float*, float) {}

// TI: operand element type (float for the FP32 driver; __half / __nv_bfloat16 when this kernel
// serves as the general-shape fallback of the tensor-core path).  H: optional 16-bit shadow.
template <bool TRANS_A, int MODE, typename TI>
__global__ void __launch_bounds__(GT)
sgemm_kernel(const TI* __restrict__ X, long ldx, const TI* __restrict__ Z, long ldz,
             float* __restrict__ C, long ldc, TI* __restrict__ H, long ldh, int M, int N, int K, int kchunk) {
    __shared__ float Xs[BK][BM + 4];
    __shared__ float Zs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    const int ty = tid / (BN / TN), tx = tid % (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int u = 0; u < TM; ++u)
#pragma unroll
        for (int v = 0; v < TN; ++v) acc[u][v] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // Z tile: BK x BN, coalesced along j
        for (int idx = tid; idx < BK * BN; idx += GT) {
            int kk = idx / BN, j = idx % BN;
            int gk = k0 + kk, gj = j0 + j;
            Zs[kk][j] = (gk < kend && gj < N) ? ldf(Z + (size_t)gk * ldz + gj) : 0.f;
        }
        if (TRANS_A) {
            for (int idx = tid; idx < BK * BM; idx += GT) {
                int kk = idx / BM, i = idx % BM;
                int gk = k0 + kk, gi = i0 + i;
                Xs[kk][i] = (gk < kend && gi < M) ? ldf(X + (size_t)gk * ldx + gi) : 0.f;
            }
        } else {
            for (int idx = tid; idx < BK * BM; idx += GT) {
                int i = idx / BK, kk = idx % BK;
                int gk = k0 + kk, gi = i0 + i;
                Xs[kk][i] = (gk < kend && gi < M) ? ldf(X + (size_t)gi * ldx + gk) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float xa[TM], zb[TN];
#pragma unroll
            for (int u = 0; u < TM; ++u) xa[u] = Xs[kk][ty * TM + u];
#pragma unroll
            for (int v = 0; v < TN; ++v) zb[v] = Zs[kk][tx * TN + v];
#pragma unroll
            for (int u = 0; u < TM; ++u)
#pragma unroll
                for (int v = 0; v < TN; ++v) acc[u][v] = fmaf(xa[u], zb[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < TM; ++u) {
        int gi = i0 + ty * TM + u;
        if (gi >= M) continue;
#pragma unroll
        for (int v = 0; v < TN; ++v) {
            int gj = j0 + tx * TN + v;
            if (gj >= N) continue;
            float* p = C + (size_t)gi * ldc + gj;
            if (MODE == 0) {
                *p = acc[u][v];
                if (sizeof(TI) == 2 && H) st16(H + (size_t)gi * ldh + gj, acc[u][v]);
            }
            else if (MODE == 1) {
                float nv = *p - acc[u][v];
                *p = nv;
                if (sizeof(TI) == 2 && H) st16(H + (size_t)gi * ldh + gj, nv);
            } else atomicAdd(p, acc[u][v]);
        }
    }
}


// ---- FP32 fast path (16-byte aligned operands, M % 4 == N % 4 == 0): 128 x 128 tiles, BK = 16, 8 x 8 per thread as a
// 2 x 2 arrangement of 4 x 4 sub-tiles (conflict-free LDS.128), global -> register prefetch of the next K slice while
// the current one is multiplied.  The 64 x 64 kernel above stays as the general-shape / 16-bit-operand fallback.
// [B200] TSQR and the FP32 driver were bound by the old kernel's ~12 TFLOP/s (4 x 4 micro-tile, scalar LDS).
constexpr int FM = 128, FN = 128, FK = 16, FLD = FM + 4;

template <bool TRANS_A, int MODE>
__global__ void __launch_bounds__(256, 2)
sgemm_fast_kernel(const float* __restrict__ X, long ldx, const float* __restrict__ Z, long ldz, float* __restrict__ C, long ldc,
                  int M, int N, int K, int kchunk) {
    __shared__ __align__(16) float Xs[2][FK][FLD];
    __shared__ __align__(16) float Zs[2][FK][FLD];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * FM, j0 = blockIdx.x * FN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0.f;
    float4 xr[2], zr[2];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int k0) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int idx = tid + t * 256;
            {   // Z / S tile: FK x FN, row k contiguous in j
                const int kk = idx >> 5, j = (idx & 31) * 4;
                const int gk = k0 + kk, gj = j0 + j;
                zr[t] = (gk < kend && gj < N) ? *reinterpret_cast<const float4*>(Z + (size_t)gk * ldz + gj) : zero4;
            }
            if (TRANS_A) {
                const int kk = idx >> 5, i = (idx & 31) * 4;
                const int gk = k0 + kk, gi = i0 + i;
                xr[t] = (gk < kend && gi < M) ? *reinterpret_cast<const float4*>(X + (size_t)gk * ldx + gi) : zero4;
            } else {
                const int i = idx >> 2, kq = (idx & 3) * 4;
                const int gk = k0 + kq, gi = i0 + i;
                xr[t] = (gk < kend && gi < M) ? *reinterpret_cast<const float4*>(X + (size_t)gi * ldx + gk) : zero4;
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int idx = tid + t * 256;
            {
                const int kk = idx >> 5, j = (idx & 31) * 4;
                *reinterpret_cast<float4*>(&Zs[buf][kk][j]) = zr[t];
            }
            if (TRANS_A) {
                const int kk = idx >> 5, i = (idx & 31) * 4;
                *reinterpret_cast<float4*>(&Xs[buf][kk][i]) = xr[t];
            } else {
                const int i = idx >> 2, kq = (idx & 3) * 4;
                Xs[buf][kq][i] = xr[t].x; Xs[buf][kq + 1][i] = xr[t].y; Xs[buf][kq + 2][i] = xr[t].z; Xs[buf][kq + 3][i] = xr[t].w;
            }
        }
    };
    fetch(kbeg);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += FK) {
        const bool more = k0 + FK < kend;
        if (more) fetch(k0 + FK);
#pragma unroll
        for (int kk = 0; kk < FK; ++kk) {
            const float4 xa = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 4]);
            const float4 xb = *reinterpret_cast<const float4*>(&Xs[buf][kk][64 + ty * 4]);
            const float4 za = *reinterpret_cast<const float4*>(&Zs[buf][kk][tx * 4]);
            const float4 zb = *reinterpret_cast<const float4*>(&Zs[buf][kk][64 + tx * 4]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float zv[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(xv[u], zv[v], acc[u][v]);
        }
        if (more) {
            stash(buf ^ 1);   // (the other buffer: its readers finished before the barrier that ended the previous slice)
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int gi = i0 + (u < 4 ? ty * 4 + u : 64 + ty * 4 + (u - 4));
        if (gi >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gj = j0 + h * 64 + tx * 4;
            if (gj >= N) continue;
            float* p = C + (size_t)gi * ldc + gj;
            const float4 a4 = make_float4(acc[u][4 * h], acc[u][4 * h + 1], acc[u][4 * h + 2], acc[u][4 * h + 3]);
            if (MODE == 0) {
                *reinterpret_cast<float4*>(p) = a4;
            } else if (MODE == 1) {
                float4 c4 = *reinterpret_cast<const float4*>(p);
                c4.x -= a4.x; c4.y -= a4.y; c4.z -= a4.z; c4.w -= a4.w;
                *reinterpret_cast<float4*>(p) = c4;
            } else {
                atomicAdd(p, a4.x); atomicAdd(p + 1, a4.y); atomicAdd(p + 2, a4.z); atomicAdd(p + 3, a4.w);
            }
        }
    }
}

inline bool fast_ok(const void* X, long ldx, const void* Z, long ldz, const void* C, long ldc, int M, int N) {
    return ((ldx | ldz | ldc) & 3) == 0 && ((M | N) & 3) == 0 && M >= 64 && N >= 64 &&
           ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(C)) & 15) == 0;
}

int tn_fast(const float* X, long ldx, const float* Z, long ldz, float* S, long lds, int M, int N, int K, cudaStream_t stream) {
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    const int tiles = ceil_div(M, FM) * ceil_div(N, FN);
    int splits = 1;
    if (K > 4 * FK) {
        const int want = ceil_div(2 * sm_count(di), tiles);   // two CTAs per SM
        const int maxs = ceil_div(K, 8 * FK);
        splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    }
    const int kchunk = round_up(ceil_div(K, splits), FK);
    splits = ceil_div(K, kchunk);
    dim3 grid(ceil_div(N, FN), ceil_div(M, FM), splits);
    if (splits > 1) {
        MPQR_CUDA(cudaMemset2DAsync(S, lds * sizeof(float), 0, (size_t)N * sizeof(float), M, stream));
        sgemm_fast_kernel<true, 2><<<grid, 256, 0, stream>>>(X, ldx, Z, ldz, S, lds, M, N, K, kchunk);
    } else {
        sgemm_fast_kernel<true, 0><<<grid, 256, 0, stream>>>(X, ldx, Z, ldz, S, lds, M, N, K, kchunk);
    }
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

int nn_fast(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K, cudaStream_t stream, int store) {
    dim3 grid(ceil_div(N, FN), ceil_div(M, FM), 1);
    if (store) sgemm_fast_kernel<false, 0><<<grid, 256, 0, stream>>>(X, ldx, S, lds, C, ldc, M, N, K, K);
    else sgemm_fast_kernel<false, 1><<<grid, 256, 0, stream>>>(X, ldx, S, lds, C, ldc, M, N, K, K);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

template <typename TI>
int tn_any(const TI* X, long ldx, const TI* Z, long ldz, float* S, long lds, int M, int N, int K,
           cudaStream_t stream) {
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    int tiles = ceil_div(M, BM) * ceil_div(N, BN);
    int splits = 1;
    if (K > 4 * BK) {
        int want = ceil_div(4 * sm_count(di), tiles);
        int maxs = ceil_div(K, 8 * BK);
        splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    }
    int kchunk = round_up(ceil_div(K, splits), BK);
    splits = ceil_div(K, kchunk);
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    if (splits > 1) {
        MPQR_CUDA(cudaMemset2DAsync(S, lds * sizeof(float), 0, (size_t)N * sizeof(float), M, stream));
        sgemm_kernel<true, 2, TI><<<grid, GT, 0, stream>>>(X, ldx, Z, ldz, S, lds, nullptr, 0, M, N, K, kchunk);
    } else {
        sgemm_kernel<true, 0, TI><<<grid, GT, 0, stream>>>(X, ldx, Z, ldz, S, lds, nullptr, 0, M, N, K, kchunk);
    }
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

template <typename TI>
int nn_any(const TI* X, long ldx, const TI* S, long lds, float* C, long ldc, TI* H, long ldh, int M, int N, int K,
           cudaStream_t stream, int store = 0) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
    if (store) sgemm_kernel<false, 0, TI><<<grid, GT, 0, stream>>>(X, ldx, S, lds, C, ldc, H, ldh, M, N, K, K);
    else sgemm_kernel<false, 1, TI><<<grid, GT, 0, stream>>>(X, ldx, S, lds, C, ldc, H, ldh, M, N, K, K);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

}  // namespace

int preload_simt_gemm() {
    const void* fns[] = {
        (const void*)sgemm_fast_kernel<true, 0>, (const void*)sgemm_fast_kernel<true, 2>, (const void*)sgemm_fast_kernel<false, 0>,
        (const void*)sgemm_fast_kernel<false, 1>,
        (const void*)sgemm_kernel<true, 0, float>, (const void*)sgemm_kernel<true, 2, float>, (const void*)sgemm_kernel<false, 0, float>,
        (const void*)sgemm_kernel<false, 1, float>,
        (const void*)sgemm_kernel<true, 0, __half>, (const void*)sgemm_kernel<true, 2, __half>, (const void*)sgemm_kernel<false, 0, __half>,
        (const void*)sgemm_kernel<false, 1, __half>,
        (const void*)sgemm_kernel<true, 0, __nv_bfloat16>, (const void*)sgemm_kernel<true, 2, __nv_bfloat16>,
        (const void*)sgemm_kernel<false, 0, __nv_bfloat16>, (const void*)sgemm_kernel<false, 1, __nv_bfloat16>};
    for (const void* f : fns) {
        cudaFuncAttributes fa;
        MPQR_CUDA(cudaFuncGetAttributes(&fa, f));
    }
    return MPQR_OK;
}

int sgemm_tn(const float* X, long ldx, const float* Z, long ldz, float* S, long lds, int M, int N,
             int K, cudaStream_t stream, long* launches) {
    if (M <= 0 || N <= 0) return MPQR_OK;
    if (fast_ok(X, ldx, Z, ldz, S, lds, M, N)) MPQR_TRY(tn_fast(X, ldx, Z, ldz, S, lds, M, N, K, stream));
    else MPQR_TRY(tn_any<float>(X, ldx, Z, ldz, S, lds, M, N, K, stream));
    if (launches) *launches += 1;
    return MPQR_OK;
}

int sgemm_nn_sub(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M,
                 int N, int K, cudaStream_t stream, long* launches) {
    if (M <= 0 || N <= 0 || K <= 0) return MPQR_OK;
    if (fast_ok(X, ldx, S, lds, C, ldc, M, N) && (K & 3) == 0) MPQR_TRY(nn_fast(X, ldx, S, lds, C, ldc, M, N, K, stream, 0));
    else MPQR_TRY(nn_any<float>(X, ldx, S, lds, C, ldc, nullptr, 0, M, N, K, stream));
    if (launches) *launches += 1;
    return MPQR_OK;
}

// C = X S (store), used by the TSQR thin-Q product
int sgemm_nn_store(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K,
                   cudaStream_t stream) {
    if (M <= 0 || N <= 0) return MPQR_OK;
    if (fast_ok(X, ldx, S, lds, C, ldc, M, N) && (K & 3) == 0) return nn_fast(X, ldx, S, lds, C, ldc, M, N, K, stream, 1);
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
    sgemm_kernel<false, 0, float><<<grid, GT, 0, stream>>>(X, ldx, S, lds, C, ldc, nullptr, 0, M, N, K, K);
    MPQR_CUDA(cudaGetLastError());
    return MPQR_OK;
}

// General-shape fallbacks of the tensor-core GEMMs (16-bit operands, FP32 accumulate on CUDA
// cores): used only when a sub-block does not start on a 16-byte boundary, which TMA requires.
int simt16_gemm_tn(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, int M, int N, int K,
                   int bf16, cudaStream_t stream) {
    if (bf16) return tn_any<__nv_bfloat16>((const __nv_bfloat16*)X, ldx, (const __nv_bfloat16*)Z, ldz, S, lds, M, N, K, stream);
    return tn_any<__half>((const __half*)X, ldx, (const __half*)Z, ldz, S, lds, M, N, K, stream);
}

int simt16_gemm_nn(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16, long ldc16,
                   int M, int N, int K, int bf16, cudaStream_t stream, int store) {
    if (bf16) return nn_any<__nv_bfloat16>((const __nv_bfloat16*)X, ldx, (const __nv_bfloat16*)S16, lds16, C, ldc, (__nv_bfloat16*)C16, ldc16, M, N, K, stream, store);
    return nn_any<__half>((const __half*)X, ldx, (const __half*)S16, lds16, C, ldc, (__half*)C16, ldc16, M, N, K, stream, store);
}

}  // namespace mpqr

// Tuning hook (tools/sgemm_time.py; not part of the public header): the FP32 GEMMs on their own.
// op 0: S = X^T Z (X: K x M, Z: K x N)   1: C -= X S (X: M x K, S: K x N)   2: C = X S
extern "C" int mpqr_debug_sgemm(int op, const float* X, long ldx, const float* Z, long ldz, float* C, long ldc, int M, int N, int K,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (op == 0) return mpqr::sgemm_tn(X, ldx, Z, ldz, C, ldc, M, N, K, st, nullptr);
    if (op == 1) return mpqr::sgemm_nn_sub(X, ldx, Z, ldz, C, ldc, M, N, K, st, nullptr);
    return mpqr::sgemm_nn_store(X, ldx, Z, ldz, C, ldc, M, N, K, st);
}
