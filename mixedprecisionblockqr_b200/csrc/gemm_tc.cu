// gemm_tc.cu — tcgen05 / TMEM / TMA GEMMs of the mixed-precision trailing update (sm_100a).
//
// These two kernels replace, on Blackwell tensor cores, the reference's trailing-matrix
// machinery: shared_mem_mmult_in_place_transpose_a (Cuda/mmult.cu:236-288, FP32 SIMT with a
// dense (m-l)x(m-l) panel-Q), dev_cpy_strided_array (Cuda/mmult.cuh:104-151), the three
// pad/cast passes dev_cpy_and_cast_array (Cuda/mmult.cuh:153-200) and the WMMA kernel
// dev_tensorcore_mmult_tiled (Cuda/mmult.cuh:252-300):
//
//   TN:  S[M x N]  = X^T Z      X:[K x M], Z:[K x N] 16-bit row-major  (contraction over ROWS;
//                               both operands are "MN-major" for the tensor core)
//   NN:  C[M x N] -= X S        X:[M x K] 16-bit row-major (K-major), S:[K x N] 16-bit
//                               row-major (MN-major); C is the FP32 master, the updated value
//                               is also mirrored into a 16-bit shadow (operand of the next TN)
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0  : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1  : MMA issuer     (one elected lane, tcgen05.mma.cta_group::1.kind::f16, FP32
//                             accumulators double-buffered in TMEM, tcgen05.commit -> mbarrier)
//   warp 2  : TMEM allocator
//   warp 3  : C-chunk loader (NN only: TMA-prefetches the FP32 master tile in 128x32 chunks)
//   warps 4-7: epilogue      (tcgen05.ld -> registers -> swizzled smem -> TMA store /
//                             TMA reduce-add for split-K)
// All global traffic goes through TMA; arbitrary (M,N,K) are handled by tensor-map bounds
// (zero fill on load, clipping on store).
#include <unordered_map>

#include "common.cuh"

namespace mpqr {
namespace {

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint64_t adesc, uint64_t bdesc, uint32_t tmem_d, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------- descriptors
// UMMA shared-memory descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: c_format F32 [4,6)=1 | a_format [7,10) | b_format [10,13) |
// a_major [15] | b_major [16] (1 = MN-major) | N>>3 [17,23) | M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int fmt16, int a_mn, int b_mn, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt16 << 7) | ((uint32_t)fmt16 << 10) | ((uint32_t)a_mn << 15) |
           ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BM = 128;      // UMMA M (TMEM lanes)
constexpr int BK = 64;       // K per pipeline stage (= one 128-byte swizzle span of 16-bit data)
constexpr int UK = 16;       // UMMA K for 16-bit operands
constexpr int CCH = 32;      // epilogue chunk: 32 fp32 columns = 128 B
constexpr int NCS = 4;       // C-chunk ring slots (16 KB each)
constexpr int NHS = 2;       // 16-bit staging slots (8 KB each)
constexpr int NTHREADS = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 3 : 4;
    static constexpr int CBUF_BYTES = BM * CCH * 4;  // 16 KB
    static constexpr int HBUF_BYTES = BM * CCH * 2;  // 8 KB
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_STAGE_BYTES;
    static constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;
    static constexpr int OFF_H = OFF_C + NCS * CBUF_BYTES;
    static constexpr int OFF_BAR = OFF_H + NHS * HBUF_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * NCS;
    static constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;  // + tmem slot + align slack
    static constexpr int TMEM_COLS = 2 * BN;
};

struct GemmParams {
    int M, N, K;
    // element coordinates of the operand blocks inside their arrays
    int ax0, ay0;  // A operand: (col, row) of its first element
    int bx0, by0;  // B operand
    int cx0, cy0;  // C / S output
    int hx0, hy0;  // 16-bit shadow
    int splits;    // split-K factor (TN only)
    int kblocks_per_split;
    int has_shadow;
    int no_c;      // TN only: the FP32 result is not wanted, only its 16-bit copy (operand of the following NN GEMM)
};

// kAMN : A operand is MN-major (TN GEMM) else K-major (NN GEMM)
// kEpi : 0 -> S = acc (store / reduce-add when splits > 1) ; 1 -> C -= acc (+ shadow) ; 2 -> C = acc (+ shadow)
template <int BN, bool kAMN, int kEpi>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmH,
               GemmParams p, int fmt16) {
    using C_ = Cfg<BN>;
    constexpr int STAGES = C_::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem + C_::OFF_A;
    uint8_t* sB = smem + C_::OFF_B;
    uint8_t* sC = smem + C_::OFF_C;
    uint8_t* sH = smem + C_::OFF_H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::OFF_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* cfull = tempty + 2;
    uint64_t* cempty = cfull + NCS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + NCS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (p.M + BM - 1) / BM, nt = (p.N + BN - 1) / BN;
    const int total = mt * nt * p.splits;
    const int kblocks = (p.K + BK - 1) / BK;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmC);
        if (p.has_shadow) prefetch_tmap(&tmH);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
        }
        for (int i = 0; i < NCS; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, C_::TMEM_COLS);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // barrier init / TMEM allocation / descriptor prefetch overlapped the predecessor's tail

    auto decode = [&](int w, int& mb, int& nb, int& kb0, int& kb1) {
        int tile = w / p.splits, ks = w - tile * p.splits;
        // consecutive CTAs share the same A (M) tile column block -> X stays L2/SMEM friendly
        mb = tile % mt;
        nb = tile / mt;
        kb0 = ks * p.kblocks_per_split;
        kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > kblocks) kb1 = kblocks;
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                int mb, nb, kb0, kb1;
                decode(w, mb, nb, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + C_::B_STAGE_BYTES);
                    uint8_t* a = sA + stage * A_STAGE_BYTES;
                    uint8_t* b = sB + stage * C_::B_STAGE_BYTES;
                    if (kAMN) {
                        // A tile: 128 (m, contiguous) x 64 (k rows): two 64x64 boxes
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_2d(a + i * (BK * 128), &tmA, &full[stage], p.ax0 + mb * BM + i * 64, p.ay0 + kb * BK);
                    } else {
                        // A tile: 64 (k, contiguous) x 128 (m rows): one box
                        tma_load_2d(a, &tmA, &full[stage], p.ax0 + kb * BK, p.ay0 + mb * BM);
                    }
                    // B tile: BN (n, contiguous) x 64 (k rows): BN/64 boxes of 64x64
#pragma unroll
                    for (int i = 0; i < BN / 64; ++i)
                        tma_load_2d(b + i * (BK * 128), &tmB, &full[stage], p.bx0 + nb * BN + i * 64, p.by0 + kb * BK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        const uint32_t idesc = make_idesc(fmt16, kAMN ? 1 : 0, 1, BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
            int mb, nb, kb0, kb1;
            decode(w, mb, nb, kb0, kb1);
            const int as = it & 1;
            mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = smem_u32(sA + stage * A_STAGE_BYTES);
                    const uint32_t b_base = smem_u32(sB + stage * C_::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        uint64_t ad, bd;
                        if (kAMN) ad = make_smem_desc(a_base + k * (UK * 128), BK * 128, 1024);
                        else ad = make_smem_desc(a_base + k * (UK * 2), 0, 1024);
                        bd = make_smem_desc(b_base + k * (UK * 128), BK * 128, 1024);
                        umma_f16(ad, bd, d_tmem, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) umma_commit(&tfull[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ================================ C-chunk loader (NN) =========================
        if (kEpi == 1 && elect_one()) {
            uint32_t g = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                int mb, nb, kb0, kb1;
                decode(w, mb, nb, kb0, kb1);
                for (int j = 0; j < BN / CCH; ++j) {
                    if (nb * BN + j * CCH >= p.N) break;
                    const int slot = g % NCS;
                    mbar_wait(&cempty[slot], ((g / NCS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&cfull[slot], C_::CBUF_BYTES);
                    tma_load_2d(sC + slot * C_::CBUF_BYTES, &tmC, &cfull[slot], p.cx0 + nb * BN + j * CCH, p.cy0 + mb * BM);
                    ++g;
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ====================================
        const int q = warp - 4;               // TMEM lane quarter
        const int row = q * 32 + lane;        // accumulator row owned by this thread
        const bool leader = (threadIdx.x == 128);
        uint32_t g = 0;
        int it = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
            int mb, nb, kb0, kb1;
            decode(w, mb, nb, kb0, kb1);
            const int as = it & 1;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            for (int j = 0; j < BN / CCH; ++j) {
                if (nb * BN + j * CCH >= p.N) break;
                const int slot = g % NCS;
                const int hs = g % NHS;
                uint8_t* cb = sC + slot * C_::CBUF_BYTES;
                uint8_t* hb = sH + hs * C_::HBUF_BYTES;
                // stores of chunks <= g-2 have finished reading smem
                if (leader) {
                    tma_wait_read<1>();
                    if (kEpi == 1 && g >= 2) mbar_arrive(&cempty[(g - 2) % NCS]);
                }
                if (kEpi == 1) mbar_wait(&cfull[slot], (g / NCS) & 1);
                named_bar_sync(1, 128);

                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + j * CCH), v);
                tmem_ld_wait();
                uint8_t* crow = cb + row * 128;
                float o[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4* pc = reinterpret_cast<float4*>(crow + ((i ^ (row & 7)) << 4));
                    float4 a4 = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    float4 r4;
                    if (kEpi == 1) {
                        float4 c4 = *pc;
                        r4 = make_float4(c4.x - a4.x, c4.y - a4.y, c4.z - a4.z, c4.w - a4.w);
                    } else {
                        r4 = a4;
                    }
                    if (kEpi != 0 || !p.no_c) *pc = r4;
                    o[4 * i] = r4.x; o[4 * i + 1] = r4.y; o[4 * i + 2] = r4.z; o[4 * i + 3] = r4.w;
                }
                if (p.has_shadow) {
                    uint8_t* hrow = hb + row * 64;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float lo = o[8 * i + 2 * u], hi = o[8 * i + 2 * u + 1];
                            if (fmt16 == 1) {
                                __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                                pk[u] = *reinterpret_cast<uint32_t*>(&t);
                            } else {
                                __half2 t = __floats2half2_rn(lo, hi);
                                pk[u] = *reinterpret_cast<uint32_t*>(&t);
                            }
                        }
                        *reinterpret_cast<uint4*>(hrow + ((i ^ ((row >> 1) & 3)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (j == BN / CCH - 1 || nb * BN + (j + 1) * CCH >= p.N) {
                    // last TMEM read of this tile: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&tempty[as]);
                }
                fence_proxy_async();
                named_bar_sync(2, 128);
                if (leader) {
                    const int cx = p.cx0 + nb * BN + j * CCH, cy = p.cy0 + mb * BM;
                    if (kEpi == 0 && p.splits > 1) tma_reduce_add_2d(&tmC, cb, cx, cy);
                    else if (kEpi != 0 || !p.no_c) tma_store_2d(&tmC, cb, cx, cy);
                    if (p.has_shadow) tma_store_2d(&tmH, hb, p.hx0 + nb * BN + j * CCH, p.hy0 + mb * BM);
                    tma_commit();
                }
                ++g;
            }
        }
        if (leader) tma_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C_::TMEM_COLS);
}

// ------------------------------------------------------------------------ 2-CTA variant (cta_group::2)
// A CTA pair (cluster of 2 = the two SMs of a TPC) computes a 256 x 256 tile: every CTA stages its own 128 rows of A and
// HALF of the B tile (128 of the 256 columns), one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256),
// which reads the B halves of both CTAs, and every CTA keeps the accumulators of its own 128 rows in its own TMEM.
// Per k-block a CTA pulls 32 KB through its L2 port instead of 48 KB: the one-CTA kernel sits exactly at the port's
// ~64 B/clk at full tensor rate (768 KB of operands per 128 x 256 x 1024 tile in 7.2 us), which is what capped TN at
// ~60 % and NN (+ the FP32 master tile) at ~42 % of the tensor peak in round 1.
//   full[]  : in the LEADER only; both CTAs' TMA loads complete their bytes there (cp.async.bulk.tensor.cta_group::2)
//   empty[] : in both CTAs, arrived by the leader's tcgen05.commit multicast
//   tfull[] : in both CTAs (multicast commit);  tempty[]: in the leader, 2 x 128 remote arrivals from both epilogues
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint64_t adesc, uint64_t bdesc, uint32_t tmem_d, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

struct Cfg2 {
    static constexpr int BN = 256;               // pair tile: 256 x 256
    static constexpr int BNH = 128;              // columns of B staged per CTA
    static constexpr int B_STAGE_BYTES = BNH * BK * 2;  // 16 KB
    static constexpr int STAGES = 4;             // (3 stages + 7 C slots measured slower on B200: TN 1382 -> 1243, NN 800 -> 708 TFLOP/s)
    static constexpr int NCS2 = 5;               // C-chunk ring slots
    static constexpr int CBUF_BYTES = BM * CCH * 4;
    static constexpr int HBUF_BYTES = BM * CCH * 2;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_STAGE_BYTES;
    static constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;
    static constexpr int OFF_H = OFF_C + NCS2 * CBUF_BYTES;
    static constexpr int OFF_BAR = OFF_H + NHS * HBUF_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * NCS2;
    static constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
};
static_assert(Cfg2::SMEM_BYTES <= 232448, "2-CTA GEMM: shared memory budget");

template <bool kAMN, int kEpi>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmH,
                GemmParams p, int fmt16) {
    using C_ = Cfg2;
    constexpr int STAGES = C_::STAGES, BN = C_::BN, NCS2 = C_::NCS2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem + C_::OFF_A;
    uint8_t* sB = smem + C_::OFF_B;
    uint8_t* sC = smem + C_::OFF_C;
    uint8_t* sH = smem + C_::OFF_H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::OFF_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* cfull = tempty + 2;
    uint64_t* cempty = cfull + NCS2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + NCS2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader_cta = rank == 0;
    const int mt = (p.M + 2 * BM - 1) / (2 * BM), nt = (p.N + BN - 1) / BN;
    const int total = mt * nt * p.splits;
    const int kblocks = (p.K + BK - 1) / BK;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmC);
        if (p.has_shadow) prefetch_tmap(&tmH);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 256);
        }
        for (int i = 0; i < NCS2; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(tmem_slot, C_::TMEM_COLS);
    pdl_launch_dependents();
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers exist before any remote signal; TMEM address visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    auto decode = [&](int w, int& mb, int& nb, int& kb0, int& kb1) {
        int tile = w / p.splits, ks = w - tile * p.splits;
        mb = tile % mt;
        nb = tile / mt;
        kb0 = ks * p.kblocks_per_split;
        kb1 = kb0 + p.kblocks_per_split;
        if (kb1 > kblocks) kb1 = kblocks;
    };

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = pair; w < total; w += npairs) {
                int mb, nb, kb0, kb1;
                decode(w, mb, nb, kb0, kb1);
                const int m0 = mb * 2 * BM + (int)rank * BM;         // this CTA's rows of A (and of C)
                const int n0 = nb * BN + (int)rank * C_::BNH;        // this CTA's half of the B tile
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader_cta) mbar_arrive_expect_tx(&full[stage], 2 * (A_STAGE_BYTES + C_::B_STAGE_BYTES));
                    uint8_t* a = sA + stage * A_STAGE_BYTES;
                    uint8_t* b = sB + stage * C_::B_STAGE_BYTES;
                    if (kAMN) {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_2d_2sm(a + i * (BK * 128), &tmA, &full[stage], p.ax0 + m0 + i * 64, p.ay0 + kb * BK);
                    } else {
                        tma_load_2d_2sm(a, &tmA, &full[stage], p.ax0 + kb * BK, p.ay0 + m0);
                    }
#pragma unroll
                    for (int i = 0; i < C_::BNH / 64; ++i)
                        tma_load_2d_2sm(b + i * (BK * 128), &tmB, &full[stage], p.bx0 + n0 + i * 64, p.by0 + kb * BK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ================
        if (leader_cta) {
            const uint32_t idesc = make_idesc(fmt16, kAMN ? 1 : 0, 1, 2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = pair; w < total; w += npairs, ++it) {
                int mb, nb, kb0, kb1;
                decode(w, mb, nb, kb0, kb1);
                const int as = it & 1;
                mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_base = smem_u32(sA + stage * A_STAGE_BYTES);
                        const uint32_t b_base = smem_u32(sB + stage * C_::B_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            uint64_t ad, bd;
                            if (kAMN) ad = make_smem_desc(a_base + k * (UK * 128), BK * 128, 1024);
                            else ad = make_smem_desc(a_base + k * (UK * 2), 0, 1024);
                            bd = make_smem_desc(b_base + k * (UK * 128), BK * 128, 1024);
                            umma2_f16(ad, bd, d_tmem, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma2_commit_both(&empty[stage]);
                        if (kb == kb1 - 1) umma2_commit_both(&tfull[as]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ================================ C-chunk loader (NN, per CTA) ================
        if (kEpi == 1 && elect_one()) {
            uint32_t g = 0;
            for (int w = pair; w < total; w += npairs) {
                int mb, nb, kb0, kb1;
                decode(w, mb, nb, kb0, kb1);
                const int m0 = mb * 2 * BM + (int)rank * BM;
                for (int j = 0; j < BN / CCH; ++j) {
                    if (nb * BN + j * CCH >= p.N) break;
                    const int slot = g % NCS2;
                    mbar_wait(&cempty[slot], ((g / NCS2) & 1) ^ 1);
                    mbar_arrive_expect_tx(&cfull[slot], C_::CBUF_BYTES);
                    tma_load_2d(sC + slot * C_::CBUF_BYTES, &tmC, &cfull[slot], p.cx0 + nb * BN + j * CCH, p.cy0 + m0);
                    ++g;
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue (both CTAs, own rows) ==============
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const bool leader = (threadIdx.x == 128);
        uint32_t g = 0;
        int it = 0;
        for (int w = pair; w < total; w += npairs, ++it) {
            int mb, nb, kb0, kb1;
            decode(w, mb, nb, kb0, kb1);
            const int m0 = mb * 2 * BM + (int)rank * BM;
            const int as = it & 1;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            for (int j = 0; j < BN / CCH; ++j) {
                if (nb * BN + j * CCH >= p.N) break;
                const int slot = g % NCS2;
                const int hs = g % NHS;
                uint8_t* cb = sC + slot * C_::CBUF_BYTES;
                uint8_t* hb = sH + hs * C_::HBUF_BYTES;
                if (leader) {
                    tma_wait_read<1>();
                    if (kEpi == 1 && g >= 2) mbar_arrive(&cempty[(g - 2) % NCS2]);
                }
                if (kEpi == 1) mbar_wait(&cfull[slot], (g / NCS2) & 1);
                named_bar_sync(1, 128);

                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + j * CCH), v);
                tmem_ld_wait();
                uint8_t* crow = cb + row * 128;
                float o[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4* pc = reinterpret_cast<float4*>(crow + ((i ^ (row & 7)) << 4));
                    float4 a4 = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    float4 r4;
                    if (kEpi == 1) {
                        float4 c4 = *pc;
                        r4 = make_float4(c4.x - a4.x, c4.y - a4.y, c4.z - a4.z, c4.w - a4.w);
                    } else {
                        r4 = a4;
                    }
                    if (kEpi != 0 || !p.no_c) *pc = r4;
                    o[4 * i] = r4.x; o[4 * i + 1] = r4.y; o[4 * i + 2] = r4.z; o[4 * i + 3] = r4.w;
                }
                if (p.has_shadow) {
                    uint8_t* hrow = hb + row * 64;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float lo = o[8 * i + 2 * u], hi = o[8 * i + 2 * u + 1];
                            if (fmt16 == 1) {
                                __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                                pk[u] = *reinterpret_cast<uint32_t*>(&t);
                            } else {
                                __half2 t = __floats2half2_rn(lo, hi);
                                pk[u] = *reinterpret_cast<uint32_t*>(&t);
                            }
                        }
                        *reinterpret_cast<uint4*>(hrow + ((i ^ ((row >> 1) & 3)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (j == BN / CCH - 1 || nb * BN + (j + 1) * CCH >= p.N) {
                    // last TMEM read of this tile: hand the accumulator back to the leader's MMA warp
                    tc_fence_before();
                    mbar_arrive_leader(&tempty[as]);
                }
                fence_proxy_async();
                named_bar_sync(2, 128);
                if (leader) {
                    const int cx = p.cx0 + nb * BN + j * CCH, cy = p.cy0 + m0;
                    if (kEpi == 0 && p.splits > 1) tma_reduce_add_2d(&tmC, cb, cx, cy);
                    else if (kEpi != 0 || !p.no_c) tma_store_2d(&tmC, cb, cx, cy);
                    if (p.has_shadow) tma_store_2d(&tmH, hb, p.hx0 + nb * BN + j * CCH, p.hy0 + m0);
                    tma_commit();
                }
                ++g;
            }
        }
        if (leader) tma_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();  // no remote arrival / multicast commit may still be in flight towards a CTA that exits
    if (warp == 2) tmem_dealloc2(tmem_base, C_::TMEM_COLS);
}

// ------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MPQR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (!ptr || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return MPQR_ECUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    *out = fn;
    return MPQR_OK;
}

// 2-D row-major tensor map: `cols` x `rows` elements visible (everything beyond is OOB: zero
// on load, clipped on store), row pitch ld elements.
// A tensor map depends only on these parameters; a factorisation repeated on the same plan and buffers (every bench step,
// every call of the cached host plan) asks for the same ~7000 maps again, and cuTensorMapEncodeTiled is a noticeable part
// of the host's issue time once the device chain is short (host issue 96 ms of a 140 ms factorisation in round 1).
struct MapKey {
    const void* base;
    long cols, rows, ld;
    int eb, bf, bc, br, swz;
    bool operator==(const MapKey& o) const {
        return base == o.base && cols == o.cols && rows == o.rows && ld == o.ld && eb == o.eb && bf == o.bf && bc == o.bc && br == o.br && swz == o.swz;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)k.cols * 0xBF58476D1CE4E5B9ull) ^ ((uint64_t)k.rows * 0x94D049BB133111EBull) ^ ((uint64_t)k.ld << 17);
        h ^= (uint64_t)(k.eb | (k.bf << 4) | (k.bc << 8) | (k.br << 18) | (k.swz << 28));
        return (size_t)(h ^ (h >> 29));
    }
};
int make_map_uncached(CUtensorMap* map, const void* base, int elem_bytes, int is_bf16, long cols, long rows, long ld,
                      int box_cols, int box_rows, CUtensorMapSwizzle swz);
int make_map(CUtensorMap* map, const void* base, int elem_bytes, int is_bf16, long cols, long rows, long ld,
             int box_cols, int box_rows, CUtensorMapSwizzle swz) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{base, cols, rows, ld, elem_bytes, is_bf16, box_cols, box_rows, (int)swz};
    auto it = cache.find(key);
    if (it != cache.end()) { *map = it->second; return MPQR_OK; }
    MPQR_TRY(make_map_uncached(map, base, elem_bytes, is_bf16, cols, rows, ld, box_cols, box_rows, swz));
    if (cache.size() > (1u << 16)) cache.clear();
    cache.emplace(key, *map);
    return MPQR_OK;
}
int make_map_uncached(CUtensorMap* map, const void* base, int elem_bytes, int is_bf16, long cols, long rows, long ld,
                      int box_cols, int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc;
    MPQR_TRY(get_encode(&enc));
    if (((uintptr_t)base & 15) || ((ld * elem_bytes) & 15)) {
        set_error("tensor map: base %p / pitch %ld B not 16-byte aligned", base, ld * elem_bytes);
        return MPQR_EINVAL;
    }
    CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : (is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (cols=%ld rows=%ld ld=%ld box=%dx%d)", (int)r, cols, rows, ld,
                  box_cols, box_rows);
        return MPQR_ECUDA;
    }
    return MPQR_OK;
}

template <int BN, bool kAMN, int kEpi>
int launch(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tH,
           const GemmParams& p, int fmt16, int grid, cudaStream_t stream) {
    using C_ = Cfg<BN>;
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<BN, kAMN, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
    cudaLaunchAttribute pat[1] = {pdl_attr()};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.stream = stream; cfg.attrs = pat; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = C_::SMEM_BYTES;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, kAMN, kEpi>, tA, tB, tC, tH, p, fmt16));
    return MPQR_OK;
}

template <bool kAMN, int kEpi>
int launch2(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tH,
            const GemmParams& p, int fmt16, int pairs, cudaStream_t stream) {
    MPQR_TRY(func_attr_once((const void*)tc_gemm2_kernel<kAMN, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1] = pdl_attr();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NTHREADS); cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 2;
    cfg.dynamicSmemBytes = Cfg2::SMEM_BYTES;
    MPQR_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<kAMN, kEpi>, tA, tB, tC, tH, p, fmt16));
    return MPQR_OK;
}
}  // namespace
int preload_tc_gemm() {
    EncodeTiledFn enc;
    MPQR_TRY(get_encode(&enc));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<256, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<128, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<256, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<128, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<256, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm_kernel<128, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm2_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
    MPQR_TRY(func_attr_once((const void*)tc_gemm2_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
    return MPQR_OK;
}
namespace {
// 2-CTA kernel: worth it from two 128-row tiles on (MPQR_GEMM_1CTA=1 keeps the one-CTA kernel: A/B comparisons)
inline bool use_2cta(int M, int N) { return M > BM && N > 128 && !getenv("MPQR_GEMM_1CTA"); }

// TMA needs 16-byte aligned box origins.  x0 = misalignment of a block pointer in elements;
// blocks with x0 != 0 are routed to the CUDA-core fallback (gemm_simt.cu).  TMA stores also
// write whole 16-byte granules (measured on B200: a store clipped at column N still zeroes the
// rest of the granule), so an output block whose last column is not granule-aligned may only go
// through TMA when the caller says the spill is harmless (pad_ok: row padding / dead columns).
struct Blk {
    const void* base;
    int x0;
};
Blk align_blk(const void* ptr, int elem_bytes) {
    uintptr_t a = (uintptr_t)ptr;
    uintptr_t b = a & ~(uintptr_t)15;
    return {(const void*)b, (int)((a - b) / elem_bytes)};
}

int pick_bn(int N) { return N > 128 ? 256 : 128; }

}  // namespace

int tc_gemm_tn(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, int M, int N, int K,
               int bf16, int pad_ok, cudaStream_t stream, long* launches) {
    return tc_gemm_tn16(X, ldx, Z, ldz, S, lds, nullptr, 0, nullptr, M, N, K, bf16, pad_ok, stream, launches);
}

// S16 != null: when the launch needs no split-K the epilogue rounds the result to 16 bit itself and writes ONLY S16
// (*wrote16 = 1; S stays untouched); otherwise S (FP32) is produced as usual and the caller converts (*wrote16 = 0).
int tc_gemm_tn16(const void* X, long ldx, const void* Z, long ldz, float* S, long lds, void* S16, long lds16, int* wrote16,
                 int M, int N, int K, int bf16, int pad_ok, cudaStream_t stream, long* launches) {
    if (wrote16) *wrote16 = 0;
    if (M <= 0 || N <= 0) return MPQR_OK;
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    const int BN = pick_bn(N);
    Blk bx = align_blk(X, 2), bz = align_blk(Z, 2), bs = align_blk(S, 4);
    if (bx.x0 || bz.x0 || bs.x0 || (ldx & 7) || (ldz & 7) || (lds & 3) || (!pad_ok && (N & 3))) {
        // TMA box origins must be 16-byte aligned: general-shape CUDA-core fallback
        int rc = simt16_gemm_tn(X, ldx, Z, ldz, S, lds, M, N, K, bf16, stream);
        if (rc == MPQR_OK && launches) *launches += 1;
        return rc;
    }
    CUtensorMap tA, tB, tC;
    MPQR_TRY(make_map(&tA, bx.base, 2, bf16, bx.x0 + M, K, ldx, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    MPQR_TRY(make_map(&tB, bz.base, 2, bf16, bz.x0 + N, K, ldz, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    MPQR_TRY(make_map(&tC, bs.base, 4, 0, bs.x0 + N, M, lds, CCH, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.ax0 = bx.x0; p.bx0 = bz.x0; p.cx0 = bs.x0;
    const bool two = use_2cta(M, N);
    const int units = two ? sm_count(di) / 2 : sm_count(di);   // CTAs, or CTA pairs
    const int tiles = two ? ceil_div(M, 2 * BM) * ceil_div(N, 256) : ceil_div(M, BM) * ceil_div(N, BN);
    const int kblocks = ceil_div(K, BK);
    int splits = 1;
    if (tiles < units && kblocks >= 8) {
        splits = units / tiles;
        int maxs = kblocks / 4;
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    p.kblocks_per_split = ceil_div(kblocks, splits);
    p.splits = ceil_div(kblocks, p.kblocks_per_split);
    CUtensorMap tH = tC;
    if (S16 && p.splits == 1 && !align_blk(S16, 2).x0 && !(lds16 & 7) && (pad_ok || !(N & 7))) {
        MPQR_TRY(make_map(&tH, S16, 2, bf16, N, M, lds16, CCH, BM, CU_TENSOR_MAP_SWIZZLE_64B));
        p.has_shadow = 1;
        p.no_c = 1;
        if (wrote16) *wrote16 = 1;
    }
    if (p.splits > 1)
        MPQR_CUDA(cudaMemset2DAsync(S, lds * sizeof(float), 0, (size_t)N * sizeof(float), M, stream));
    int total = tiles * p.splits;
    int grid = total < units ? total : units;
    int rc = two ? launch2<true, 0>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream)
                 : (BN == 256) ? launch<256, true, 0>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream)
                               : launch<128, true, 0>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream);
    if (rc == MPQR_OK && launches) *launches += 1;
    return rc;
}

static int tc_gemm_nn_impl(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16, long ldc16,
                           int M, int N, int K, int bf16, int pad_ok, int store, cudaStream_t stream, long* launches) {
    if (M <= 0 || N <= 0 || K <= 0) return MPQR_OK;
    DeviceInfo di;
    MPQR_TRY(get_device_info(&di));
    const int BN = pick_bn(N);
    Blk bx = align_blk(X, 2), bs = align_blk(S16, 2), bc = align_blk(C, 4);
    if (bx.x0 || bs.x0 || bc.x0 || (C16 && align_blk(C16, 2).x0) || (ldx & 7) || (lds16 & 7) || (ldc & 3) ||
        (C16 && (ldc16 & 7)) || (!pad_ok && (N & (C16 ? 7 : 3)))) {
        int rc = simt16_gemm_nn(X, ldx, S16, lds16, C, ldc, C16, ldc16, M, N, K, bf16, stream, store);
        if (rc == MPQR_OK && launches) *launches += 1;
        return rc;
    }
    CUtensorMap tA, tB, tC, tH;
    MPQR_TRY(make_map(&tA, bx.base, 2, bf16, bx.x0 + K, M, ldx, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    MPQR_TRY(make_map(&tB, bs.base, 2, bf16, bs.x0 + N, K, lds16, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    MPQR_TRY(make_map(&tC, bc.base, 4, 0, bc.x0 + N, M, ldc, CCH, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.ax0 = bx.x0; p.bx0 = bs.x0; p.cx0 = bc.x0;
    p.splits = 1;
    p.kblocks_per_split = ceil_div(K, BK);
    if (C16) {
        Blk bh = align_blk(C16, 2);
        MPQR_TRY(make_map(&tH, bh.base, 2, bf16, bh.x0 + N, M, ldc16, CCH, BM, CU_TENSOR_MAP_SWIZZLE_64B));
        p.hx0 = bh.x0;
        p.has_shadow = 1;
    } else {
        tH = tC;
    }
    const bool two = use_2cta(M, N);
    const int units = two ? sm_count(di) / 2 : sm_count(di);
    const int tiles = two ? ceil_div(M, 2 * BM) * ceil_div(N, 256) : ceil_div(M, BM) * ceil_div(N, BN);
    int grid = tiles < units ? tiles : units;
    int rc;
    if (two) rc = store ? launch2<false, 2>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream)
                        : launch2<false, 1>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream);
    else if (store) rc = (BN == 256) ? launch<256, false, 2>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream)
                                : launch<128, false, 2>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream);
    else rc = (BN == 256) ? launch<256, false, 1>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream)
                          : launch<128, false, 1>(tA, tB, tC, tH, p, bf16 ? 1 : 0, grid, stream);
    if (rc == MPQR_OK && launches) *launches += 1;
    return rc;
}

int tc_gemm_nn(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16, long ldc16,
               int M, int N, int K, int bf16, int pad_ok, cudaStream_t stream, long* launches) {
    return tc_gemm_nn_impl(X, ldx, S16, lds16, C, ldc, C16, ldc16, M, N, K, bf16, pad_ok, 0, stream, launches);
}

int tc_gemm_nn_store(const void* X, long ldx, const void* S16, long lds16, float* C, long ldc, void* C16, long ldc16,
                     int M, int N, int K, int bf16, cudaStream_t stream, long* launches) {
    return tc_gemm_nn_impl(X, ldx, S16, lds16, C, ldc, C16, ldc16, M, N, K, bf16, 1, 1, stream, launches);
}

}  // namespace mpqr
