// greenctx_probe.cu — can two green contexts (16 SMs cluster-capable + the rest) run kernels concurrently,
// and does a 16-CTA cluster launch work inside the small partition?  nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include <algorithm>
#define CK(x) do { CUresult r__ = (x); if (r__ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r__, &s); printf("FAIL %s -> %s\n", #x, s); return 1; } } while (0)
#define RK(x) do { cudaError_t r__ = (x); if (r__ != cudaSuccess) { printf("FAIL %s -> %s\n", #x, cudaGetErrorString(r__)); return 1; } } while (0)

__global__ void spin_kernel(long long cycles, int* smids) {
    if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = (int)s; }
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
__global__ void cluster_kernel(long long cycles, int* smids) {
    if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = (int)s; }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}

int main() {
    RK(cudaFree(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUdevResource full; CK(cuDeviceGetDevResource(dev, &full, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", full.sm.smCount);
    for (unsigned want : {16u, 24u, 32u}) {
        CUdevResource grp[1], rem; unsigned n = 1;
        CUresult r = cuDevSmResourceSplitByCount(grp, &n, &full, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE, want);
        const char* s = "ok"; if (r != CUDA_SUCCESS) cuGetErrorString(r, &s);
        printf("split minCount=%u (cluster flag): %s groups=%u group0=%u remaining=%u\n", want, s, n, r == CUDA_SUCCESS ? grp[0].sm.smCount : 0, r == CUDA_SUCCESS ? rem.sm.smCount : 0);
    }
    CUdevResource grp[1], rem; unsigned n = 1;
    CK(cuDevSmResourceSplitByCount(grp, &n, &full, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE, 16));
    CUdevResourceDesc dA, dB;
    CK(cuDevResourceGenerateDesc(&dA, &grp[0], 1));
    CK(cuDevResourceGenerateDesc(&dB, &rem, 1));
    CUgreenCtx gA, gB;
    CK(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CK(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream sA, sB;
    CK(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
    CK(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
    int *smA, *smB;
    RK(cudaMalloc(&smA, 4096 * 4)); RK(cudaMalloc(&smB, 4096 * 4));
    RK(cudaMemset(smA, 0xff, 4096 * 4)); RK(cudaMemset(smB, 0xff, 4096 * 4));
    cudaEvent_t e0, e1, a0, a1;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&a0); cudaEventCreate(&a1);
    RK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const long long ms1 = 1900000;  // ~1 ms
    // B: one long kernel (20 ms) with 2 CTAs per SM worth of grid; A: 100 cluster-16 kernels of ~0.05 ms each
    RK(cudaDeviceSynchronize());
    RK(cudaEventRecord(e0, (cudaStream_t)sB));
    spin_kernel<<<296, 128, 0, (cudaStream_t)sB>>>(10 * ms1, smB);
    RK(cudaGetLastError());
    RK(cudaEventRecord(e1, (cudaStream_t)sB));
    RK(cudaEventRecord(a0, (cudaStream_t)sA));
    for (int i = 0; i < 100; ++i) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(16); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)sA;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        RK(cudaLaunchKernelEx(&cfg, cluster_kernel, ms1 / 20, smA));
    }
    RK(cudaEventRecord(a1, (cudaStream_t)sA));
    RK(cudaDeviceSynchronize());
    float tb, ta, gap;
    cudaEventElapsedTime(&tb, e0, e1); cudaEventElapsedTime(&ta, a0, a1); cudaEventElapsedTime(&gap, e0, a1);
    printf("B (296 CTAs x 10 ms spin on the big partition): %.2f ms ; A (100 cluster-16 launches x 0.05 ms): %.2f ms ; B start -> A end: %.2f ms\n", tb, ta, gap);
    std::vector<int> ha(16), hb(296);
    cudaMemcpy(ha.data(), smA, 16 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), smB, 296 * 4, cudaMemcpyDeviceToHost);
    std::sort(ha.begin(), ha.end()); std::sort(hb.begin(), hb.end());
    hb.erase(std::unique(hb.begin(), hb.end()), hb.end());
    printf("A smids:"); for (int v : ha) printf(" %d", v); printf("\nB distinct smids: %zu (min %d max %d)\n", hb.size(), hb.front(), hb.back());
    int overlap = 0; for (int v : ha) if (std::binary_search(hb.begin(), hb.end(), v)) ++overlap;
    printf("SMs used by both: %d\n", overlap);
    // runtime-API kernel with <<<>>> on the primary context's default stream still works?
    spin_kernel<<<148, 128>>>(1000, smB);
    RK(cudaDeviceSynchronize());
    printf("OK\n");
    return 0;
}
