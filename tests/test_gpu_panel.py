"""Panel kernel (Householder vectors + fused compact-WY) vs the oracle restatement of
h_householder_qr (Cuda/qr.cu:198-293) and h_wy_transform (Cuda/qr.cu:337-426).
Tolerances: FP32 rounding with different (tree vs sequential) summation order."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _check_panel(m, n, lam, pw, seed, tol=2e-5):
    from gpu_util import panel_factor
    A = oracle.uniform_matrix(m, n, seed)
    P0 = oracle.pack(A)
    if lam:
        P0 = oracle.householder_panel(P0.copy(), 0, lam)  # realistic state: earlier columns factored
    Pref = oracle.householder_panel(P0.copy(), lam, pw)
    Wref, Yref = oracle.wy_factors(Pref, lam, pw)
    P, Y, W, T = panel_factor(P0.copy(), lam, pw)
    scale = np.abs(Pref).max()
    assert np.abs(P - Pref).max() <= tol * scale, np.abs(P - Pref).max()
    # untouched outside the panel columns
    mask = np.ones_like(P, bool)
    mask[lam:, lam:lam + pw] = False
    assert np.array_equal(P[mask], P0[mask])
    assert np.abs(Y - Yref).max() <= tol
    assert np.abs(W - Wref).max() <= 20 * tol, np.abs(W - Wref).max()
    # T: W = Y T
    assert np.abs(Y.astype(np.float64) @ T.astype(np.float64) - W).max() <= 20 * tol
    assert np.allclose(np.tril(T, -1), 0)


@pytest.mark.parametrize("m,n,lam,pw", [
    (6, 4, 0, 2), (12, 8, 0, 5), (12, 8, 5, 3), (60, 40, 16, 16), (97, 90, 80, 10), (129, 80, 64, 16),
    (300, 64, 0, 64), (600, 400, 128, 32), (2048, 128, 0, 128), (2048, 2048, 0, 32), (5000, 256, 128, 128),
    (4096, 100, 32, 68),
])
def test_panel_vs_oracle(m, n, lam, pw):
    _check_panel(m, n, lam, pw, seed=m * 31 + n)


def test_panel_large_multi_cta():
    # taller than one cluster can hold in registers: shared-memory fallback kernel
    _check_panel(40000, 128, 0, 128, seed=3, tol=5e-5)


@pytest.mark.parametrize("m,n,lam,pw", [
    (9000, 128, 0, 128),      # 32-wide register blocks, cluster of 16 + 2 rows/thread
    (20000, 128, 0, 128),     # 16-wide register blocks, 4 rows/thread
    (32768, 64, 0, 64),       # capacity limit of one cluster
    (3000, 96, 32, 50),       # ragged block widths (32 + 18)
    (700, 40, 0, 40),
])
def test_panel_register_blocks_tall(m, n, lam, pw):
    _check_panel(m, n, lam, pw, seed=m + pw, tol=5e-5)


def test_panel_zero_column_and_signs():
    # zero column is skipped (Cuda/qr.cu:242-244); sign(0) = +1 (:229-235); known 3x3 answer (:1397-1401)
    from gpu_util import panel_factor
    A = np.array([[12, -51, 4], [6, 167, -68], [-4, 24, -41]], np.float32)
    P, _, _, _ = panel_factor(oracle.pack(A), 0, 3)
    gold = np.array([[-14, -21, 14], [.9636241, -175, 70], [.2223748, .9984604, 35], [-.1482499, .0554700, -1]], np.float32)
    assert np.allclose(P, gold, atol=2e-5)
    Z = np.array([[0, 3, 1], [0, 4, -2], [0, 1, 1]], np.float32)  # first column all zero
    P, Y, W, T = panel_factor(oracle.pack(Z), 0, 3)
    Pref = oracle.householder_panel(oracle.pack(Z), 0, 3)
    assert np.allclose(P, Pref, atol=1e-5)
    assert np.all(Y[:, 0] == 0) and np.all(W[:, 0] == 0)
    # [0,0,2] -> v = [1,0,1]/sqrt2, image [-2,0,0]  (python/test_all.py:12-20)
    V = np.array([[0.0], [0.0], [2.0]], np.float32)
    P, _, _, _ = panel_factor(oracle.pack(V), 0, 1)
    assert np.allclose(P[:, 0], [-2, 1 / np.sqrt(2), 0, 1 / np.sqrt(2)], atol=1e-6)


@pytest.mark.parametrize("m,n,lam,pw", [(2048, 128, 0, 128), (20000, 128, 0, 128), (5000, 256, 128, 128), (700, 64, 0, 64)])
@pytest.mark.parametrize("mode", ["classic", "gate"])
def test_panel_flows_agree(m, n, lam, pw, mode, monkeypatch):
    # the persistent chain in its default ordering (side updates behind cuStreamWaitValue32, side flag posted by the U kernel)
    # is covered by the tests above; here the per-block launch flow it replaces (MPQR_NO_CHAIN=1 -- what MPQR_STREAM_ORDERED
    # handles, odd widths and unaligned shapes take) and the chain with one-thread gate kernels instead of stream memory
    # operations (MPQR_GATE_KERNEL=1: the fallback for drivers without the entry point)
    monkeypatch.setenv({"classic": "MPQR_NO_CHAIN", "gate": "MPQR_GATE_KERNEL"}[mode], "1")
    _check_panel(m, n, lam, pw, seed=m + 7 * pw, tol=5e-5)
