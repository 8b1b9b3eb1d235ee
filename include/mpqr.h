/*
 * mpqr.h — C-ABI of libmpqr.so: B200-native (sm_100a) mixed-precision blocked Householder QR.
 *
 * Drop-in boundary for the block-QR path of jaidonlybbert/MixedPrecisionBlockQR.  The
 * reference has no FFI layer; its operator API is four C++ free functions with
 * C-compatible signatures (reference Cuda/qr.cuh:129-137):
 *     void dev_mixed_precision_block_qr(float* A, float* Q, int m, int n, int r);
 *     void dev_block_qr_wy            (float* A, float* Q, int m, int n, int r);
 *     void dev_block_qr               (float* A, float* Q, int m, int n, int r);
 *     void h_block_qr                 (float* A, float* Q, int m, int n, int r);
 * Every entry point below names the reference interface it replaces.  Plain pointers and
 * sizes only; no C++ / torch types.  All matrices are ROW-MAJOR FP32.
 *
 * Packed factor layout (reference Cuda/qr.cu:283-285, :1062):
 *   buffer of (m+1) rows x n cols; on return rows<=cols hold R, and the UNIT Householder
 *   vector w_k (H_k = I - 2 w_k w_k^T) of column k sits at rows k+1..m of column k (one
 *   row BELOW the diagonal).  sign rule w ~ u + sign(u0)||u|| e1 with sign(0)=+1; an
 *   all-zero column is skipped.
 *
 * Error behaviour: the reference is `void` and exit(1)s on CUDA errors
 * (Cuda/helper_cuda.h:583-595).  Here every call returns 0 on success or a negative
 * MPQR_E* code and never exits; mpqr_last_error() gives the message.  There is NO CPU
 * fallback: without a CUDA device the calls fail with MPQR_ECUDA.
 */
#ifndef MPQR_H_
#define MPQR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPQR_OK 0
#define MPQR_EINVAL (-1)  /* bad argument */
#define MPQR_ECUDA (-2)   /* CUDA runtime / driver error (see mpqr_last_error) */
#define MPQR_ENOMEM (-3)  /* device allocation failed */
#define MPQR_ENCCL (-4)   /* NCCL error */
#define MPQR_ESTATE (-5)  /* call sequence error (e.g. form_q before factor) */

/* flags */
#define MPQR_FP32 0x0u       /* FP32 SIMT trailing update  (replaces dev_block_qr_wy, Cuda/qr.cu:958) */
#define MPQR_FP16 0x1u       /* FP16 operands / FP32 accumulate on tcgen05 (replaces
                                dev_mixed_precision_block_qr, Cuda/qr.cu:1049) */
#define MPQR_BF16 0x2u       /* BF16 operands / FP32 accumulate on tcgen05 */
#define MPQR_PRECISION_MASK 0x3u
#define MPQR_KEEP_WY 0x10u   /* retain W (=Y T) for every outer block so that Q can be formed /
                                WY factors read back after the factorisation */
#define MPQR_STREAM_ORDERED 0x20u /* panels as plain stream-ordered kernels: no persistent panel kernel that waits on device
                                flags for work issued behind it.  Use it when several handles factor concurrently on one device
                                next to other host-side CUDA activity: [B200] the TSQR driver, 4 lanes with persistent kernels
                                plus a plan being built meanwhile, stopped for good in about one run of three (csrc/tsqr.cu);
                                a single handle was never seen to (tools/alloc_hazard.py allocates while it runs).  Costs
                                ~15 % of the panel time of tall panels. */

typedef struct mpqr_handle mpqr_handle;

const char* mpqr_last_error(void);
const char* mpqr_version(void);

/* ---------------------------------------------------------------------------------------
 * Host-pointer drop-ins (same ownership contract as the reference drivers: caller owns
 * A ((m+1)*n floats, rows 0..m-1 = input, row m = 0) and Q (m*m floats); callee moves
 * data to the current device, factors, moves results back, leaves nothing resident).
 *   mpqr_block_qr_host(..., MPQR_FP16) == dev_mixed_precision_block_qr  (Cuda/qr.cu:1049-1226)
 *   mpqr_block_qr_host(..., MPQR_FP32) == dev_block_qr_wy / dev_block_qr (Cuda/qr.cu:958, :877)
 * Q may be NULL (skip explicit Q).  On entry Q's contents are ignored (the reference
 * requires identity, Cuda/qr.cu:1868-1872; identity in => same result).
 * r is the reference's panel width; internally min(r,128) columns are factored per panel
 * (the packed result does not depend on the grouping beyond rounding).
 * ------------------------------------------------------------------------------------- */
int mpqr_block_qr_host(float* A_packed, float* Q, int m, int n, int r, unsigned flags);
/* One difference from the reference's "leaves nothing resident" (Cuda/qr.cu:1221-1226 frees everything): the plan of
 * the LAST host call (workspaces + device copies of A and Q, ~12 GB at 32768^2) is kept for the next call with the
 * same shape, because allocating and freeing it costs more than the factorisation.  mpqr_release_cache() frees it
 * (and the TSQR plans); the environment variable MPQR_NO_HOST_CACHE=1 restores allocate-and-free per call. */
int mpqr_release_cache(void);

/* ---------------------------------------------------------------------------------------
 * Device-resident API (what bench.py times as `value`): plan once, factor many times.
 * ------------------------------------------------------------------------------------- */
/* Allocates all workspaces for an m x n problem on the current device.
 * r  : panel width (1..128 effective).  nb : outer block width for two-level blocking
 * (multiple of the effective r; 0 = automatic).  flags as above. */
int mpqr_create(mpqr_handle** out, int m, int n, int r, int nb, unsigned flags);
int mpqr_destroy(mpqr_handle* h);

/* Factor in place.  dA: device pointer to the packed (m+1) x lda FP32 buffer (lda >= n,
 * row m must be present; its content on entry is ignored).  stream: cudaStream_t (NULL =
 * legacy default stream).  Asynchronous with respect to the host.
 * Replaces the body of dev_mixed_precision_block_qr's panel loop, Cuda/qr.cu:1074-1219. */
int mpqr_factor_device(mpqr_handle* h, float* dA, long lda, void* stream);

/* Explicit Q = H_1 ... H_min(m,n) (m x m, ldq >= m) by blocked backward accumulation
 * (reference: Q[:, l:] <- Q[:, l:] (I - W Y^T), Cuda/qr.cu:1109-1207; h_q_backward_accumulation
 * Cuda/qr.cu:296-335).  Needs MPQR_KEEP_WY at create time and a preceding factor call. */
int mpqr_form_q_device(mpqr_handle* h, float* dQ, long ldq, void* stream);

/* Compact-WY / WY factors of the last factorisation (north_star: "returning Q, or the WY
 * factors, and R").  Panel p covers columns [p*r_eff, min((p+1)*r_eff, kmax)).
 * T: r_eff x r_eff upper triangular FP32 with Q_p = I - Y_p T_p Y_p^T, copied to dT (device,
 * ld >= r_eff).  Replaces what dev_wy_transform (Cuda/qr.cu:535-600) computes and frees. */
int mpqr_get_panel_T(mpqr_handle* h, int panel, float* dT, int ldt, void* stream);
int mpqr_num_panels(const mpqr_handle* h);
int mpqr_effective_r(const mpqr_handle* h);
int mpqr_effective_nb(const mpqr_handle* h);
/* Number of kernels launched by the last factor / form_q call (for bench.py's gpu_launches). */
long mpqr_last_launch_count(const mpqr_handle* h);

/* Per-kernel-class timing with CUDA events recorded on the launch stream (bench.py's roofline
 * object).  Classes: 0 = panel factorisation (+fused WY), 1 = GEMM TN (S = W^T A), 2 = GEMM NN
 * (A -= Y S, incl. shadow), 3 = casts / small utility kernels.  Enable before a factor call;
 * mpqr_get_profile synchronises on the recorded events and returns totals accumulated since the
 * last mpqr_set_profiling(h, 1): device milliseconds, launches, algorithmic flops and
 * algorithmic HBM bytes (panel: 8*D*pw; NN: 10*M*N; TN: 2*K*(M+N)+4*M*N). */
#define MPQR_NUM_KERNEL_CLASSES 8  /* 4..7: parts of class 0 (register-block kernels, in-panel S, Gram/T/W, in-panel U) */
int mpqr_set_profiling(mpqr_handle* h, int on);
int mpqr_get_profile(mpqr_handle* h, int kernel_class, double* ms_total, long* launches, double* flops,
                     double* bytes);

/* ---------------------------------------------------------------------------------------
 * Single kernels exposed for parity tests against the oracle (tests/ call these through
 * ctypes).  All pointers are device pointers.
 * ------------------------------------------------------------------------------------- */
/* Panel factorisation of columns [lam, lam+pw) rows [lam, m) of the packed buffer; also
 * emits Y, W = Y T (D x pw row-major FP32, ld = pw, D = m-lam) and T (pw x pw).
 * Replaces h_householder_qr (Cuda/qr.cu:198-293) + dev_wy_transform's W/Y (Cuda/qr.cu:535-600).
 * Any of dY, dW, dT may be NULL. */
int mpqr_panel_factor_device(float* dA, long lda, int m, int n, int lam, int pw,
                             float* dY, float* dW, float* dT, void* stream);

/* Tensor-core GEMM primitives of the trailing update (tcgen05, FP16/BF16 operands, FP32
 * accumulate); they replace shared_mem_mmult_in_place_transpose_a + dev_cpy_strided_array
 * (Cuda/mmult.cu:236-288, Cuda/mmult.cuh:104-151) and dev_tensorcore_mmult_tiled
 * (Cuda/mmult.cuh:252-300).  `bf16` selects the operand type.  Operands are 16-bit,
 * row-major, leading dimensions multiples of 8 elements, base pointers 16-byte aligned.
 *   tn : S[M x N] (fp32, lds) = X^T Z,  X is [K x M] (ldx), Z is [K x N] (ldz)
 *   nn : C[M x N] (fp32, ldc) -= X S,   X is [M x K] (ldx), S is [K x N] (lds16);
 *        if dC16 != NULL the updated C is also written rounded to 16 bit (ldc16). */
int mpqr_gemm_tn_device(const void* dX, long ldx, const void* dZ, long ldz, float* dS, long lds,
                        int M, int N, int K, int bf16, void* stream);
int mpqr_gemm_nn_device(const void* dX, long ldx, const void* dS16, long lds16, float* dC, long ldc,
                        void* dC16, long ldc16, int M, int N, int K, int bf16, void* stream);

/* Deterministic synthetic input: element (i,j) of an m x n matrix = uniform[0,1) from a
 * stateless hash of (seed, i*n+j) (distribution of h_generate_random_matrix,
 * Cuda/mmult.cuh:39-60).  Writes rows [row0,row0+rows) x cols [col0,col0+cols) into dA
 * (row-major, lda), i.e. any shard can be generated in place on any GPU. */
int mpqr_fill_uniform_device(float* dA, long lda, long n_total, long row0, long rows, long col0,
                             long cols, uint64_t seed, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU): 1-D column-block-cyclic over `nranks` GPUs of one
 * NVSwitch box; each panel block is factored by its owner and its Y/W broadcast with NCCL.
 * The NCCL communicator is owned by the library; bootstrap the unique id through any
 * out-of-band channel (bench.py uses torch.distributed).
 * ------------------------------------------------------------------------------------- */
#define MPQR_NCCL_UID_BYTES 128
/* Pure host layout helpers (no device needed): columns of an n-column matrix owned by `rank`
 * for block width nb, and the global column of a local column (-1 if out of range). */
int mpqr_mg_layout_local_cols(int n, int nb, int rank, int nranks);
int mpqr_mg_layout_global_col(int n, int nb, int rank, int nranks, int local_col);
int mpqr_mg_get_unique_id(void* uid_out /* MPQR_NCCL_UID_BYTES */);
int mpqr_mg_create(mpqr_handle** out, int m, int n, int r, int nb, unsigned flags, int rank,
                   int nranks, const void* uid);
/* Number of columns this rank owns and the global column of local column j. */
int mpqr_mg_local_cols(const mpqr_handle* h);
int mpqr_mg_global_col(const mpqr_handle* h, int local_col);
/* dA_local: (m+1) x lda_local packed buffer holding this rank's columns (block-cyclic). */
int mpqr_mg_factor_device(mpqr_handle* h, float* dA_local, long lda_local, void* stream);

/* ---------------------------------------------------------------------------------------
 * Least-squares solve on top of a factorisation: the reference's dev_QR_Solver
 * (Cuda/QR/Solver/solver.cu:39-87) is an unimplemented stub of GVL 5.3.2, x = R^-1 Q^T b;
 * python/linear_least_sqare.py:5-22 is its NumPy demo.  dA_packed: the factor produced by
 * mpqr_factor_device on THIS handle (m >= n).  dB: m x nrhs row-major (ldb >= nrhs, 1 <= nrhs <= 8);
 * on return rows 0..n-1 hold the solution x, rows n..m-1 the remaining components of Q^T b
 * (their norm is the residual).  FP32 arithmetic.
 * ------------------------------------------------------------------------------------- */
int mpqr_solve_device(mpqr_handle* h, const float* dA_packed, long lda, float* dB, long ldb, int nrhs, void* stream);

/* ---------------------------------------------------------------------------------------
 * EuRoC Jacobian text file -> packed (rows+1) x cols host buffer (replaces read_euroc_jacobian,
 * Cuda/qr.cu:696-776: "<rows> <cols>" then 0-based "<row> <col> <value>" triples, zero fill).
 * Host only.  Free the buffer with mpqr_free_host.
 * ------------------------------------------------------------------------------------- */
int mpqr_read_euroc_jacobian(const char* path, int* rows, int* cols, float** packed_out);
void mpqr_free_host(void* p);

/* ---------------------------------------------------------------------------------------
 * Device-side parity metrics: the reference's harness metrics are O(m^3) FP32 host
 * loops (Cuda/qr.cu:85-196); these compute the same quantities on the GPU with FP64 accumulation,
 * never materialising Q R or Q^T Q.  All matrix pointers are DEVICE pointers (row-major FP32),
 * results are written to HOST doubles and the calls synchronise `stream`.  Deterministic.
 *   mpqr_strip_r_device               h_strip_R_from_A         (Cuda/qr.cu:85-100)
 *   mpqr_backward_error_device        h_backward_error         (Cuda/qr.cu:115-135): ||A0 - Q R||_F / ||A0||_F;
 *                                     dR is read through the mask row <= col, so the packed factor may be
 *                                     passed directly; a_norm (may be NULL) receives ||A0||_F
 *   mpqr_q_error_device               h_q_error                (Cuda/qr.cu:137-171): max_signed is the
 *                                     reference's quantity (max entry of Q^T Q - I, no abs), max_abs and
 *                                     fro = ||Q^T Q - I||_F are north_star's orthogonality figures (any may be NULL)
 *   mpqr_lower_trapezoid_error_device h_lower_trapezoid_error  (Cuda/qr.cu:173-196): ||strict lower part||_F
 *   mpqr_frobenius_norm_device        h_matrix_norm
 *   mpqr_r_agreement_device           north_star's "elementwise |R| agreement": max | |R| - |Rref| | over
 *                                     row <= col, max |Rref|, and the Frobenius norm of the difference
 * ------------------------------------------------------------------------------------- */
int mpqr_strip_r_device(const float* dA_packed, long lda, float* dR, long ldr, int m, int n, void* stream);
int mpqr_backward_error_device(const float* dA0, long lda0, const float* dR, long ldr, const float* dQ, long ldq,
                               int m, int n, double* err, double* a_norm, void* stream);
int mpqr_q_error_device(const float* dQ, long ldq, int m, double* max_signed, double* max_abs, double* fro, void* stream);
int mpqr_lower_trapezoid_error_device(const float* dR, long ldr, int m, int n, double* err, void* stream);
int mpqr_frobenius_norm_device(const float* dX, long ldx, long rows, long cols, double* nrm, void* stream);
int mpqr_r_agreement_device(const float* dR, long ldr, const float* dRref, long ldref, int m, int n,
                            double* max_abs_diff, double* max_abs_ref, double* fro_diff, void* stream);
/* Host: the reference's own operation count 4m^2n - mn^2 + n^3/3 per second (h_qr_flops_per_second,
 * Cuda/qr.cu:102-113) and its result log (h_write_results_to_log, Cuda/qr.cu:58-83): appends
 * "rows,cols,runtime,flops,error" lines to <log_dir>/<file_name>.txt, the format
 * Cuda/performance/util.py:19-31 reads.  log_dir NULL = "log", file_name NULL = "logFile" (the
 * reference's defaults). */
float mpqr_qr_flops_per_second(float time_ms, int m, int n);
int mpqr_write_results_to_log(const char* log_dir, const char* file_name, int height, int width, float time_ms,
                              float flops_per_second, float backward_error);

/* ---------------------------------------------------------------------------------------
 * Tall-skinny QR (replaces python/ca_qr.py:25-43 ts_qr): A is m x n (m >> n), row-major FP32
 * on the device.  Row blocks are factored independently, the n x n R factors are reduced by
 * a tree; R (n x n, ldr) is returned and, if dQ != NULL, the thin Q (m x n, ldq).  The call is SYNCHRONOUS with
 * respect to the host (it returns after `stream` and its internal lanes have finished).
 * ------------------------------------------------------------------------------------- */
int mpqr_tsqr_device(const float* dA, long lda, long m, int n, float* dQ, long ldq, float* dR,
                     long ldr, void* stream);
/* mpqr_tsqr_device keeps its plans (streams, workspaces; ~0.7 GB for 1048576 x 256) for the next call with
 * the same shape; this frees them. */
int mpqr_tsqr_release_cache(void);

/* Multi-GPU TSQR (BASELINE config 5): 1-D ROW-block layout, one process per GPU, rank p holds
 * m_local rows of A (row-major, device).  Local TSQR, ONE ncclAllGather of the n x n R factors, the
 * (nranks*n) x n stack factored redundantly on every rank (ts_qr's tree, python/ca_qr.py:36-41, with nranks
 * leaves); rank 0's R is then broadcast, so dR (n x n) is identical on all ranks.  dQ_local (m_local x n, may be NULL) receives this rank's
 * rows of the thin Q.  The handle only owns the NCCL communicator (destroy with mpqr_destroy). */
int mpqr_mg_tsqr_create(mpqr_handle** out, int rank, int nranks, const void* uid);
int mpqr_mg_tsqr_device(mpqr_handle* h, const float* dA_local, long lda, long m_local, int n, float* dQ_local,
                        long ldq, float* dR, long ldr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPQR_H_ */
