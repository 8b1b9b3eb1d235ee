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

int sgemm_tn(const float* X, long ldx, const float* Z, long ldz, float* S, long lds, int M, int N,
             int K, cudaStream_t stream, long* launches) {
    if (M <= 0 || N <= 0) return MPQR_OK;
    MPQR_TRY(tn_any<float>(X, ldx, Z, ldz, S, lds, M, N, K, stream));
    if (launches) *launches += 1;
    return MPQR_OK;
}

int sgemm_nn_sub(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M,
                 int N, int K, cudaStream_t stream, long* launches) {
    if (M <= 0 || N <= 0 || K <= 0) return MPQR_OK;
    MPQR_TRY(nn_any<float>(X, ldx, S, lds, C, ldc, nullptr, 0, M, N, K, stream));
    if (launches) *launches += 1;
    return MPQR_OK;
}

// C = X S (store), used by the TSQR thin-Q product
int sgemm_nn_store(const float* X, long ldx, const float* S, long lds, float* C, long ldc, int M, int N, int K,
                   cudaStream_t stream) {
    if (M <= 0 || N <= 0) return MPQR_OK;
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
