"""mixedprecisionblockqr_b200 — B200-native mixed-precision blocked Householder QR.

Thin Python host mirror over the C-ABI of libmpqr.so (include/mpqr.h).  The functions
`dev_mixed_precision_block_qr`, `dev_block_qr_wy`, `dev_block_qr` keep the names, argument
meaning and in-place semantics of the reference drivers (reference Cuda/qr.cuh:129-137,
Cuda/qr.cu:877/:958/:1049) so parity tests read like the reference's own tests.

There is NO CPU fallback: every compute entry point goes through libmpqr.so and raises
MpqrError when the library or a CUDA device is missing.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpqr.so")

MPQR_FP32 = 0x0
MPQR_FP16 = 0x1
MPQR_BF16 = 0x2
MPQR_KEEP_WY = 0x10
MPQR_STREAM_ORDERED = 0x20   # no persistent (flag-waiting) panel kernel: see include/mpqr.h
NCCL_UID_BYTES = 128

# every symbol include/mpqr.h declares (tests/test_abi.py checks they are all exported)
ABI_SYMBOLS = (
    "mpqr_last_error", "mpqr_version", "mpqr_block_qr_host", "mpqr_create", "mpqr_destroy",
    "mpqr_factor_device", "mpqr_form_q_device", "mpqr_get_panel_T", "mpqr_num_panels",
    "mpqr_effective_r", "mpqr_effective_nb", "mpqr_last_launch_count", "mpqr_set_profiling", "mpqr_get_profile",
    "mpqr_panel_factor_device",
    "mpqr_gemm_tn_device", "mpqr_gemm_nn_device", "mpqr_fill_uniform_device", "mpqr_mg_layout_local_cols",
    "mpqr_mg_layout_global_col", "mpqr_mg_get_unique_id",
    "mpqr_mg_create", "mpqr_mg_local_cols", "mpqr_mg_global_col", "mpqr_mg_factor_device",
    "mpqr_tsqr_device", "mpqr_solve_device", "mpqr_read_euroc_jacobian", "mpqr_free_host",
    "mpqr_strip_r_device", "mpqr_backward_error_device", "mpqr_q_error_device", "mpqr_lower_trapezoid_error_device",
    "mpqr_frobenius_norm_device", "mpqr_r_agreement_device", "mpqr_qr_flops_per_second", "mpqr_write_results_to_log",
    "mpqr_mg_tsqr_create", "mpqr_mg_tsqr_device", "mpqr_tsqr_release_cache", "mpqr_release_cache",
)


class MpqrError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libmpqr.so; fails loudly (no fallback) if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpqrError(f"{LIB_PATH} is missing: run `python -m mixedprecisionblockqr_b200.build` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        vp, c_int, c_long, c_uint = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_uint
        L.mpqr_last_error.restype = ctypes.c_char_p
        L.mpqr_version.restype = ctypes.c_char_p
        L.mpqr_block_qr_host.argtypes = [vp, vp, c_int, c_int, c_int, c_uint]
        L.mpqr_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, c_int, c_int, c_uint]
        L.mpqr_destroy.argtypes = [vp]
        L.mpqr_factor_device.argtypes = [vp, vp, c_long, vp]
        L.mpqr_form_q_device.argtypes = [vp, vp, c_long, vp]
        L.mpqr_get_panel_T.argtypes = [vp, c_int, vp, c_int, vp]
        for f in ("mpqr_num_panels", "mpqr_effective_r", "mpqr_effective_nb", "mpqr_mg_local_cols"):
            getattr(L, f).argtypes = [vp]
        L.mpqr_last_launch_count.argtypes = [vp]
        L.mpqr_last_launch_count.restype = c_long
        L.mpqr_set_profiling.argtypes = [vp, c_int]
        L.mpqr_get_profile.argtypes = [vp, c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_long),
                                       ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.mpqr_panel_factor_device.argtypes = [vp, c_long, c_int, c_int, c_int, c_int, vp, vp, vp, vp]
        L.mpqr_gemm_tn_device.argtypes = [vp, c_long, vp, c_long, vp, c_long, c_int, c_int, c_int, c_int, vp]
        L.mpqr_gemm_nn_device.argtypes = [vp, c_long, vp, c_long, vp, c_long, vp, c_long, c_int, c_int, c_int, c_int, vp]
        L.mpqr_fill_uniform_device.argtypes = [vp, c_long, c_long, c_long, c_long, c_long, c_long, ctypes.c_uint64, vp]
        L.mpqr_mg_layout_local_cols.argtypes = [c_int] * 4
        L.mpqr_mg_layout_global_col.argtypes = [c_int] * 5
        L.mpqr_mg_get_unique_id.argtypes = [vp]
        L.mpqr_mg_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, c_int, c_int, c_uint, c_int, c_int, vp]
        L.mpqr_mg_global_col.argtypes = [vp, c_int]
        L.mpqr_mg_factor_device.argtypes = [vp, vp, c_long, vp]
        L.mpqr_tsqr_device.argtypes = [vp, c_long, c_long, c_int, vp, c_long, vp, c_long, vp]
        L.mpqr_solve_device.argtypes = [vp, vp, c_long, vp, c_long, c_int, vp]
        L.mpqr_read_euroc_jacobian.argtypes = [ctypes.c_char_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(ctypes.POINTER(ctypes.c_float))]
        L.mpqr_free_host.argtypes = [vp]
        dp = ctypes.POINTER(ctypes.c_double)
        L.mpqr_strip_r_device.argtypes = [vp, c_long, vp, c_long, c_int, c_int, vp]
        L.mpqr_backward_error_device.argtypes = [vp, c_long, vp, c_long, vp, c_long, c_int, c_int, dp, dp, vp]
        L.mpqr_q_error_device.argtypes = [vp, c_long, c_int, dp, dp, dp, vp]
        L.mpqr_lower_trapezoid_error_device.argtypes = [vp, c_long, c_int, c_int, dp, vp]
        L.mpqr_frobenius_norm_device.argtypes = [vp, c_long, c_long, c_long, dp, vp]
        L.mpqr_r_agreement_device.argtypes = [vp, c_long, vp, c_long, c_int, c_int, dp, dp, dp, vp]
        L.mpqr_qr_flops_per_second.argtypes = [ctypes.c_float, c_int, c_int]
        L.mpqr_qr_flops_per_second.restype = ctypes.c_float
        L.mpqr_write_results_to_log.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_int, c_int, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_float]
        L.mpqr_free_host.restype = None
        L.mpqr_mg_tsqr_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp]
        L.mpqr_mg_tsqr_device.argtypes = [vp, vp, c_long, c_long, c_int, vp, c_long, vp, c_long, vp]
        _lib = L
    return _lib


def check(rc, what="mpqr call"):
    if rc != 0:
        raise MpqrError(f"{what} failed (code {rc}): {lib().mpqr_last_error().decode()}")


def householder_flops(m, n):
    """BASELINE.json count 2mn^2 - 2n^3/3 (m >= n); 2m^2 n - 2m^3/3 for m < n."""
    m, n = float(m), float(n)
    return 2 * m * n * n - 2 * n ** 3 / 3 if m >= n else 2 * m * m * n - 2 * m ** 3 / 3


# ------------------------------------------------------------------ reference-named host drivers
def _host_driver(A, Q, m, n, r, flags):
    if not (isinstance(A, np.ndarray) and A.dtype == np.float32 and A.flags.c_contiguous and A.size == (m + 1) * n):
        raise MpqrError("A must be a C-contiguous float32 array of (m+1)*n elements (reference Cuda/qr.cu:1866)")
    if Q is not None and not (Q.dtype == np.float32 and Q.flags.c_contiguous and Q.size == m * m):
        raise MpqrError("Q must be a C-contiguous float32 array of m*m elements")
    rc = lib().mpqr_block_qr_host(A.ctypes.data, Q.ctypes.data if Q is not None else None, m, n, r, flags)
    check(rc, "mpqr_block_qr_host")


def dev_mixed_precision_block_qr(A, Q, m, n, r, bf16=False):
    """In-place drop-in for reference dev_mixed_precision_block_qr (Cuda/qr.cu:1049): A is the
    (m+1) x n packed host buffer, Q the m x m host buffer (or None)."""
    _host_driver(A, Q, m, n, r, MPQR_BF16 if bf16 else MPQR_FP16)


def dev_block_qr_wy(A, Q, m, n, r):
    """In-place drop-in for reference dev_block_qr_wy (Cuda/qr.cu:958), FP32 trailing update."""
    _host_driver(A, Q, m, n, r, MPQR_FP32)


dev_block_qr = dev_block_qr_wy  # reference Cuda/qr.cu:877 (older variant, same contract)


# ------------------------------------------------------------------ device-resident plan
class BlockQR:
    """Device-resident plan (mpqr_create / mpqr_factor_device / mpqr_form_q_device).  Pointers
    are raw device addresses (e.g. torch.Tensor.data_ptr()); stream is a cudaStream_t int."""

    def __init__(self, m, n, r, nb=0, precision="fp16", keep_wy=False, stream_ordered=False):
        flags = {"fp32": MPQR_FP32, "fp16": MPQR_FP16, "bf16": MPQR_BF16}[precision]
        if keep_wy:
            flags |= MPQR_KEEP_WY
        if stream_ordered:
            flags |= MPQR_STREAM_ORDERED
        self._h = ctypes.c_void_p()
        check(lib().mpqr_create(ctypes.byref(self._h), m, n, r, nb, flags), "mpqr_create")
        self.m, self.n = m, n
        self.r = lib().mpqr_effective_r(self._h)
        self.nb = lib().mpqr_effective_nb(self._h)
        self.num_panels = lib().mpqr_num_panels(self._h)

    def factor(self, dA_ptr, lda, stream=0):
        check(lib().mpqr_factor_device(self._h, dA_ptr, lda, stream), "mpqr_factor_device")

    def form_q(self, dQ_ptr, ldq, stream=0):
        check(lib().mpqr_form_q_device(self._h, dQ_ptr, ldq, stream), "mpqr_form_q_device")

    def solve(self, dA_ptr, lda, dB_ptr, ldb, nrhs, stream=0):
        """Least squares x = R^-1 Q^T b on the factor this plan produced (mpqr_solve_device)."""
        check(lib().mpqr_solve_device(self._h, dA_ptr, lda, dB_ptr, ldb, nrhs, stream), "mpqr_solve_device")

    def panel_T(self, panel, dT_ptr, ldt, stream=0):
        check(lib().mpqr_get_panel_T(self._h, panel, dT_ptr, ldt, stream), "mpqr_get_panel_T")

    # classes 4..7 are parts of class 0 ("panel" = register-block kernels + in-panel updates + Gram/T/W)
    KERNEL_CLASSES = ("panel", "gemm_tn", "gemm_nn", "cast", "panel_block", "panel_s", "panel_gtw", "panel_u")

    def set_profiling(self, on):
        check(lib().mpqr_set_profiling(self._h, int(on)), "mpqr_set_profiling")

    def profile(self):
        """{class: dict(ms, launches, flops, bytes)} accumulated since set_profiling(True)."""
        out = {}
        for c, name in enumerate(self.KERNEL_CLASSES):
            ms, fl, by, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
            check(lib().mpqr_get_profile(self._h, c, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(fl), ctypes.byref(by)))
            out[name] = {"ms": ms.value, "launches": cnt.value, "flops": fl.value, "bytes": by.value}
        return out

    @property
    def last_launches(self):
        return int(lib().mpqr_last_launch_count(self._h))

    def close(self):
        if self._h:
            lib().mpqr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fill_uniform(dA_ptr, lda, n_total, row0, rows, col0, cols, seed, stream=0):
    check(lib().mpqr_fill_uniform_device(dA_ptr, lda, n_total, row0, rows, col0, cols, seed, stream), "mpqr_fill_uniform_device")


def read_euroc_jacobian(path):
    """EuRoC Jacobian text file (reference Cuda/qr.cu:696-776) -> packed (rows+1) x cols float32 numpy array."""
    m, n = ctypes.c_int(), ctypes.c_int()
    buf = ctypes.POINTER(ctypes.c_float)()
    check(lib().mpqr_read_euroc_jacobian(os.fsencode(path), ctypes.byref(m), ctypes.byref(n), ctypes.byref(buf)),
          "mpqr_read_euroc_jacobian")
    try:
        return np.ctypeslib.as_array(buf, shape=(m.value + 1, n.value)).copy()
    finally:
        lib().mpqr_free_host(buf)


# ------------------------------------------------------------------ device metrics (reference Cuda/qr.cu:58-196)
def strip_R(dA_ptr, lda, dR_ptr, ldr, m, n, stream=0):
    """h_strip_R_from_A (Cuda/qr.cu:85-100) on the device."""
    check(lib().mpqr_strip_r_device(dA_ptr, lda, dR_ptr, ldr, m, n, stream), "mpqr_strip_r_device")


def backward_error(dA0_ptr, lda0, dR_ptr, ldr, dQ_ptr, ldq, m, n, stream=0):
    """h_backward_error (Cuda/qr.cu:115-135): (||A0 - QR||_F / ||A0||_F, ||A0||_F), FP64 accumulation."""
    err, an = ctypes.c_double(), ctypes.c_double()
    check(lib().mpqr_backward_error_device(dA0_ptr, lda0, dR_ptr, ldr, dQ_ptr, ldq, m, n, ctypes.byref(err),
                                           ctypes.byref(an), stream), "mpqr_backward_error_device")
    return err.value, an.value


def q_error(dQ_ptr, ldq, m, stream=0):
    """h_q_error (Cuda/qr.cu:137-171): dict(max_signed=<the reference's quantity>, max_abs, fro) of Q^T Q - I."""
    a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    check(lib().mpqr_q_error_device(dQ_ptr, ldq, m, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), stream),
          "mpqr_q_error_device")
    return {"max_signed": a.value, "max_abs": b.value, "fro": c.value}


def lower_trapezoid_error(dR_ptr, ldr, m, n, stream=0):
    """h_lower_trapezoid_error (Cuda/qr.cu:173-196)."""
    e = ctypes.c_double()
    check(lib().mpqr_lower_trapezoid_error_device(dR_ptr, ldr, m, n, ctypes.byref(e), stream),
          "mpqr_lower_trapezoid_error_device")
    return e.value


def frobenius_norm(dX_ptr, ldx, rows, cols, stream=0):
    e = ctypes.c_double()
    check(lib().mpqr_frobenius_norm_device(dX_ptr, ldx, rows, cols, ctypes.byref(e), stream), "mpqr_frobenius_norm_device")
    return e.value


def r_agreement(dR_ptr, ldr, dRref_ptr, ldref, m, n, stream=0):
    """max | |R| - |Rref| | over row <= col, max |Rref|, ||.||_F of the difference (north_star's |R| criterion)."""
    a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    check(lib().mpqr_r_agreement_device(dR_ptr, ldr, dRref_ptr, ldref, m, n, ctypes.byref(a), ctypes.byref(b),
                                        ctypes.byref(c), stream), "mpqr_r_agreement_device")
    return {"max_abs_diff": a.value, "max_abs_ref": b.value, "fro_diff": c.value}


def qr_flops_per_second(time_ms, m, n):
    """h_qr_flops_per_second (Cuda/qr.cu:102-113): the reference's own count 4m^2n - mn^2 + n^3/3."""
    return float(lib().mpqr_qr_flops_per_second(time_ms, m, n))


def write_results_to_log(height, width, time_ms, flops_per_second, backward_error, file_name="logFile", log_dir="log"):
    """h_write_results_to_log (Cuda/qr.cu:58-83); readable by the reference's Cuda/performance/util.py."""
    check(lib().mpqr_write_results_to_log(os.fsencode(log_dir), os.fsencode(file_name), height, width, time_ms,
                                          flops_per_second, backward_error), "mpqr_write_results_to_log")


def tsqr(dA_ptr, lda, m, n, dQ_ptr, ldq, dR_ptr, ldr, stream=0):
    """Device TSQR (replaces python/ca_qr.py:25-43 ts_qr): R (n x n) and optionally thin Q (m x n)."""
    check(lib().mpqr_tsqr_device(dA_ptr, lda, m, n, dQ_ptr, ldq, dR_ptr, ldr, stream), "mpqr_tsqr_device")


# ------------------------------------------------------------------ multi-GPU (one process per GPU)
def mg_layout_local_cols(n, nb, rank, nranks):
    return lib().mpqr_mg_layout_local_cols(n, nb, rank, nranks)


def mg_layout_global_cols(n, nb, rank, nranks):
    """Global column index of every local column of `rank` (numpy int array) — pure host logic."""
    nloc = mg_layout_local_cols(n, nb, rank, nranks)
    return np.array([lib().mpqr_mg_layout_global_col(n, nb, rank, nranks, j) for j in range(nloc)], dtype=np.int64)


def mg_unique_id():
    buf = ctypes.create_string_buffer(NCCL_UID_BYTES)
    check(lib().mpqr_mg_get_unique_id(buf), "mpqr_mg_get_unique_id")
    return buf.raw


class MultiGpuBlockQR:
    """1-D column-block-cyclic plan (mpqr_mg_create / mpqr_mg_factor_device).  `uid` is the 128-byte
    NCCL unique id produced by rank 0's mg_unique_id() and shared out of band."""

    def __init__(self, m, n, r, nb, rank, nranks, uid, precision="fp16"):
        flags = {"fp16": MPQR_FP16, "bf16": MPQR_BF16}[precision]
        self._h = ctypes.c_void_p()
        check(lib().mpqr_mg_create(ctypes.byref(self._h), m, n, r, nb, flags, rank, nranks, uid), "mpqr_mg_create")
        self.m, self.n, self.rank, self.nranks = m, n, rank, nranks
        self.r = lib().mpqr_effective_r(self._h)
        self.nb = lib().mpqr_effective_nb(self._h)
        self.local_cols = lib().mpqr_mg_local_cols(self._h)

    def factor(self, dA_local_ptr, lda_local, stream=0):
        check(lib().mpqr_mg_factor_device(self._h, dA_local_ptr, lda_local, stream), "mpqr_mg_factor_device")

    set_profiling = BlockQR.set_profiling
    profile = BlockQR.profile
    KERNEL_CLASSES = BlockQR.KERNEL_CLASSES

    @property
    def last_launches(self):
        return int(lib().mpqr_last_launch_count(self._h))

    def close(self):
        if self._h:
            lib().mpqr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuTSQR:
    """Row-block TSQR over the GPUs of one box (mpqr_mg_tsqr_*; replaces python/ca_qr.py:25-43 at scale):
    rank p passes its m_local x n rows; R is returned on every rank, the thin Q rows stay local."""

    def __init__(self, rank, nranks, uid):
        self._h = ctypes.c_void_p()
        check(lib().mpqr_mg_tsqr_create(ctypes.byref(self._h), rank, nranks, uid), "mpqr_mg_tsqr_create")
        self.rank, self.nranks = rank, nranks

    def factor(self, dA_local_ptr, lda, m_local, n, dQ_local_ptr, ldq, dR_ptr, ldr, stream=0):
        check(lib().mpqr_mg_tsqr_device(self._h, dA_local_ptr, lda, m_local, n, dQ_local_ptr, ldq, dR_ptr, ldr, stream),
              "mpqr_mg_tsqr_device")

    close = MultiGpuBlockQR.close
    __del__ = MultiGpuBlockQR.__del__
