"""oracle — TEST INFRASTRUCTURE ONLY.

ctypes bindings for the CPU restatement (oracle/mpqr_oracle.c -> libmpqr_oracle.so) and,
when it has been built, the UNMODIFIED reference (oracle/_ref/libref_qr.so, built by
oracle/build_ref.sh from /root/reference/Cuda/{qr.cu,mmult.cu}).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package (mixedprecisionblockqr_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_fp = ctypes.POINTER(ctypes.c_float)
_dp = ctypes.POINTER(ctypes.c_double)


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_fp)


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def build(force=False):
    so = os.path.join(_HERE, "libmpqr_oracle.so")
    src = os.path.join(_HERE, "mpqr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_uniform01.restype = ctypes.c_float
        L.orc_uniform01.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_fill_uniform.argtypes = [_fp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_uint64]
        for name in ("orc_backward_error", "orc_backward_error_packed", "orc_q_error_max",
                     "orc_orthogonality_fro", "orc_ref_flop_model", "orc_householder_flops"):
            getattr(L, name).restype = ctypes.c_double
        L.orc_householder_flops.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_tsqr.argtypes = [_dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib = L
    return _lib


_ref = None


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_qr.so"))


def ref():
    """The unmodified reference compiled into oracle/_ref (None-safe: raises if absent)."""
    global _ref
    if _ref is None:
        L = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_qr.so"))
        for name in ("ref_h_backward_error", "ref_h_q_error", "ref_h_lower_trapezoid_error",
                     "ref_h_qr_flops_per_second"):
            getattr(L, name).restype = ctypes.c_float
        L.ref_h_qr_flops_per_second.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.c_int]
        _ref = L
    return _ref


# ----------------------------------------------------------------------------- inputs
def uniform_matrix(m, n, seed):
    """m x n float32, element (i,j) = orc_uniform01(seed, i*n+j) — vectorised numpy twin of
    oracle/mpqr_oracle.c:orc_uniform01 (bit-identical; tests/test_oracle.py checks)."""
    def mix(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        idx = np.arange(m * n, dtype=np.uint64)
        h = mix(mix(np.array([seed], dtype=np.uint64)) + idx)
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(m, n)


def pack(A):
    """(m+1) x n zero-padded packed buffer the reference drivers take (Cuda/qr.cu:1866-1875)."""
    m, n = A.shape
    P = np.zeros((m + 1, n), np.float32)
    P[:m] = A
    return P


# ----------------------------------------------------------------------------- oracle calls
def householder_panel(P, off, pw):
    m, n = P.shape[0] - 1, P.shape[1]
    lib().orc_householder_panel(_f(P), m, n, off, pw)
    return P


def wy_transform(P, off, pw, dense=False):
    m, n = P.shape[0] - 1, P.shape[1]
    D = m - off
    W = np.zeros((D, pw), np.float32)
    Y = np.zeros((D, pw), np.float32)
    Pd = np.zeros((D, D), np.float32) if dense else None
    lib().orc_wy_transform(_f(P), m, n, off, pw, _f(W), _f(Y), _f(Pd) if dense else None)
    return (W, Y, Pd) if dense else (W, Y)


def wy_factors(P, off, pw):
    m, n = P.shape[0] - 1, P.shape[1]
    D = m - off
    W = np.zeros((D, pw), np.float32)
    Y = np.zeros((D, pw), np.float32)
    lib().orc_wy_factors(_f(P), m, n, off, pw, _f(W), _f(Y))
    return W, Y


def block_qr(A, r, want_q=True, dense=False):
    """Returns (packed (m+1) x n, Q m x m or None)."""
    A = np.ascontiguousarray(A, np.float32)
    m, n = A.shape
    P = pack(A)
    Q = np.eye(m, dtype=np.float32) if (want_q or dense) else None
    if dense:
        lib().orc_block_qr_dense(_f(P), _f(Q), m, n, r)
    else:
        lib().orc_block_qr(_f(P), _f(Q) if Q is not None else None, m, n, r)
    return P, Q


def q_backward_accumulation(P):
    m, n = P.shape[0] - 1, P.shape[1]
    Q = np.zeros((m, m), np.float32)
    lib().orc_q_backward_accumulation(_f(P), _f(Q), m, n)
    return Q


def strip_R(P):
    m, n = P.shape[0] - 1, P.shape[1]
    R = np.zeros((m, n), np.float32)
    lib().orc_strip_R(_f(P), _f(R), m, n)
    return R


def backward_error(A, R, Q):
    m, n = A.shape
    return lib().orc_backward_error(_f(np.ascontiguousarray(A, np.float32)), _f(R), _f(Q), m, n)


def backward_error_packed(A, P):
    m, n = A.shape
    return lib().orc_backward_error_packed(_f(np.ascontiguousarray(A, np.float32)), _f(P), m, n)


def q_error_max(Q):
    return lib().orc_q_error_max(_f(Q), Q.shape[0])


def orthogonality_fro(Q):
    return lib().orc_orthogonality_fro(_f(Q), Q.shape[0])


def householder_flops(m, n):
    return lib().orc_householder_flops(float(m), float(n))


def tsqr(A, nblk=4):
    A = np.ascontiguousarray(A, np.float64)
    m, n = A.shape
    h = m // nblk
    Q = np.zeros((h * nblk, n))
    R = np.zeros((n, n))
    rc = lib().orc_tsqr(_d(A), m, n, nblk, _d(Q), _d(R))
    if rc != 0:
        raise ValueError("orc_tsqr: needs m//nblk >= n and nblk a power of two")
    return Q, R


# ----------------------------------------------------------------------------- reference calls
def ref_block_qr(A, r):
    A = np.ascontiguousarray(A, np.float32)
    m, n = A.shape
    P = pack(A)
    Q = np.eye(m, dtype=np.float32)
    ref().ref_h_block_qr(_f(P), _f(Q), m, n, r)
    return P, Q


def ref_householder_panel(P, off, pw):
    m, n = P.shape[0] - 1, P.shape[1]
    ref().ref_h_householder_qr(_f(P), m, n, off, pw)
    return P


def ref_wy_dense(P, off, pw):
    m, n = P.shape[0] - 1, P.shape[1]
    D = m - off
    out = np.zeros((D, D), np.float32)
    ref().ref_h_wy_transform(_f(P), _f(out), m, n, off, pw)
    return out


def ref_q_backward_accumulation(P):
    m, n = P.shape[0] - 1, P.shape[1]
    Q = np.zeros((m, m), np.float32)
    ref().ref_h_q_backward_accumulation(_f(P), _f(Q), m, n)
    return Q


def ref_dev_block_qr(A, r, mixed=True):
    """Reference GPU drivers (Cuda/qr.cu:958, :1049).  Needs a GPU."""
    A = np.ascontiguousarray(A, np.float32)
    m, n = A.shape
    P = pack(A)
    Q = np.eye(m, dtype=np.float32)
    fn = ref().ref_dev_mixed_precision_block_qr if mixed else ref().ref_dev_block_qr_wy
    fn(_f(P), _f(Q), m, n, r)
    # The reference does not check its kernel launches: a failed launch leaves its error code pending in the SHARED CUDA
    # runtime (libref_qr.so links libcudart.so like torch does), where the next torch call would trip over it.
    global last_ref_cuda_error
    last_ref_cuda_error = 0
    try:
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaDeviceSynchronize()
        last_ref_cuda_error = int(rt.cudaGetLastError())
    except OSError:
        pass
    return P, Q


last_ref_cuda_error = 0
