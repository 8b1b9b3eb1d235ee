"""tcgen05 GEMM primitives vs exact integer arithmetic (bit-exact: small-integer operands are
exact in FP16/BF16 and every partial sum is exact in FP32).  Mirrors the reference's GEMM unit
tests with analytic known answers (Cuda/mmult.cu:313-361 `120*j`, :653-705 `496*j`;
Cuda/mmult.cuh:387-435 random vs h_mmult)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ints(rng, shape, lo=-3, hi=4):
    return rng.integers(lo, hi, size=shape).astype(np.float32)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 256), (32, 100, 200), (16, 16, 16), (200, 300, 1000),
                                   (256, 512, 4096), (128, 64, 20000), (1024, 1024, 2048),
                                   (384, 640, 512), (130, 300, 8192), (1024, 4000, 16384)])   # CTA pairs with a half / mostly empty second CTA, split-K
@pytest.mark.parametrize("bf16", [False, True])
def test_gemm_tn_exact(M, N, K, bf16):
    from gpu_util import gemm_tn
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    X, Z = _ints(rng, (K, M)), _ints(rng, (K, N))
    S, pad = gemm_tn(X, Z, bf16=bf16)
    ref = X.astype(np.float64).T @ Z.astype(np.float64)
    assert np.array_equal(S.astype(np.float64), ref)
    assert np.all(pad == 7.0)  # nothing written beyond N


def test_gemm_tn_known_answer():
    # reference known answer (Cuda/mmult.cu:356): A = all ones (16x16), B[k][j] = j * k -> c = 120 j
    from gpu_util import gemm_tn
    X = np.ones((16, 16), np.float32)
    Z = np.outer(np.arange(16), np.arange(16)).astype(np.float32)
    S, _ = gemm_tn(X, Z)
    assert np.array_equal(S, np.tile(120.0 * np.arange(16, dtype=np.float32), (16, 1)))


@pytest.mark.parametrize("xoff,zoff", [(0, 8), (3, 5), (8, 1)])
def test_gemm_tn_unaligned_block(xoff, zoff):
    from gpu_util import gemm_tn
    rng = np.random.default_rng(5)
    X, Z = _ints(rng, (300, 96)), _ints(rng, (300, 130))
    S, _ = gemm_tn(X, Z, xoff=xoff, zoff=zoff)
    assert np.array_equal(S.astype(np.float64), X.astype(np.float64).T @ Z.astype(np.float64))


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (100, 70, 24), (300, 513, 32), (1000, 384, 128),
                                   (2048, 2048, 1024), (384, 300, 64), (129, 257, 1024), (4096, 1000, 512)])
@pytest.mark.parametrize("bf16", [False, True])
def test_gemm_nn_exact(M, N, K, bf16):
    from gpu_util import gemm_nn
    rng = np.random.default_rng(M + N + K)
    X, S, C = _ints(rng, (M, K), -2, 3), _ints(rng, (K, N), -2, 3), _ints(rng, (M, N), -8, 9)
    out, sh = gemm_nn(X, S, C, bf16=bf16)
    ref = C.astype(np.float64) - X.astype(np.float64) @ S.astype(np.float64)
    assert np.array_equal(out[:, :N].astype(np.float64), ref)
    assert np.all(out[:, N:] == 3.0)
    if not bf16 or np.abs(ref).max() <= 256:  # small ints exact in bf16 only up to 256
        assert np.array_equal(sh[:, :N].astype(np.float64), ref)
    assert np.all(sh[:, N:] == 5.0)


def test_gemm_nn_unaligned_c():
    from gpu_util import gemm_nn
    rng = np.random.default_rng(9)
    X, S, C = _ints(rng, (260, 48), -2, 3), _ints(rng, (48, 90), -2, 3), _ints(rng, (260, 90), -8, 9)
    out, sh = gemm_nn(X, S, C, coff=4)
    ref = C.astype(np.float64) - X.astype(np.float64) @ S.astype(np.float64)
    assert np.array_equal(out[:, 4:94].astype(np.float64), ref)
    assert np.all(out[:, :4] == 3.0) and np.all(out[:, 94:] == 3.0)
    assert np.array_equal(sh[:, 4:94].astype(np.float64), ref)


def test_gemm_one_cta_kernels_still_exact(monkeypatch):
    # M > 128 and N > 128 normally take the CTA-pair kernel (tcgen05.mma.cta_group::2); the one-CTA kernel stays in use
    # for narrower shapes and is kept covered at large ones here
    from gpu_util import gemm_nn, gemm_tn
    monkeypatch.setenv("MPQR_GEMM_1CTA", "1")
    rng = np.random.default_rng(77)
    X, Z = _ints(rng, (2048, 1024)), _ints(rng, (2048, 1024))
    S, _ = gemm_tn(X, Z)
    assert np.array_equal(S.astype(np.float64), X.astype(np.float64).T @ Z.astype(np.float64))
    X, S2, C = _ints(rng, (2048, 256), -2, 3), _ints(rng, (256, 1024), -2, 3), _ints(rng, (2048, 1024), -8, 9)
    out, sh = gemm_nn(X, S2, C)
    ref = C.astype(np.float64) - X.astype(np.float64) @ S2.astype(np.float64)
    assert np.array_equal(out[:, :1024].astype(np.float64), ref) and np.array_equal(sh[:, :1024].astype(np.float64), ref)
