"""CPU: pins the oracle (oracle/mpqr_oracle.c) to the reference's own known answers, to golden
vectors produced by the UNMODIFIED reference (tests/golden/ref_block_qr.npz, generator
tests/golden/make_golden.py) and — when oracle/_ref is present — to the reference itself, bit for bit."""
import os

import numpy as np
import pytest

import oracle
from conftest import REF_SHAPES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_block_qr.npz"))
CASES = sorted({k.split("_")[0] for k in GOLD.files if "x" in k.split("_")[0] and k.split("_")[0][0].isdigit()})


def _case(key):
    m, rest = key.split("x")
    n, r = rest.split("r")
    return int(m), int(n), int(r)


def test_generator_twins_agree():
    A = oracle.uniform_matrix(37, 29, 99)
    B = np.zeros((37, 29), np.float32)
    oracle.lib().orc_fill_uniform(B.ctypes.data_as(oracle._fp), 37, 29, 29, 99)
    assert np.array_equal(A, B)
    assert 0.0 <= A.min() and A.max() < 1.0
    assert abs(A.mean() - 0.5) < 0.05


def test_known_3x3_answer():
    # Cuda/qr.cu:1397-1401 / python/test_data.py:18-22; expected packed result from SURVEY 4 (reference run, r=2)
    A = np.array([[12, -51, 4], [6, 167, -68], [-4, 24, -41]], np.float32)
    P, Q = oracle.block_qr(A, 2, dense=True)
    assert np.array_equal(P, GOLD["known3x3_packed"]) and np.array_equal(Q, GOLD["known3x3_Q"])
    gold = np.array([[-14, -21, 14], [.9636241, -175, 70], [.2223748, .9984604, 35], [-.1482499, .0554700, -1]], np.float32)
    assert np.allclose(P, gold, atol=2e-5)
    assert np.allclose(Q, [[-.857143, .394286, -.331429], [-.428571, -.902857, .034286], [.285714, -.171429, -.942857]], atol=2e-6)


def test_reflector_known_answer():
    # python/test_all.py:12-20: [0,0,2] -> v = [1,0,1]/sqrt2, image [-2,0,0]
    P = oracle.householder_panel(oracle.pack(np.array([[0.0], [0.0], [2.0]], np.float32)), 0, 1)
    assert np.allclose(P[:, 0], [-2, 1 / np.sqrt(2), 0, 1 / np.sqrt(2)], atol=1e-7)


@pytest.mark.parametrize("key", CASES)
def test_oracle_matches_reference_golden_bit_exact(key):
    m, n, r = _case(key)
    A = oracle.uniform_matrix(m, n, int(GOLD[key + "_seed"][0]))
    P, Q = oracle.block_qr(A, r, dense=True)                    # literal restatement of h_block_qr
    assert np.array_equal(P, GOLD[key + "_packed"]) and np.array_equal(Q, GOLD[key + "_Q"])
    P1 = oracle.householder_panel(oracle.pack(A), 0, r)          # h_householder_qr
    assert np.array_equal(P1, GOLD[key + "_panel0"])
    W, Y, dense = oracle.wy_transform(P1.copy(), 0, r, dense=True)  # h_wy_transform
    assert np.array_equal(dense, GOLD[key + "_wydense0"])
    Pf = oracle.householder_panel(oracle.pack(A), 0, n)
    assert np.array_equal(Pf, GOLD[key + "_hhfull"])
    assert np.array_equal(oracle.q_backward_accumulation(Pf), GOLD[key + "_qback"])  # h_q_backward_accumulation


@pytest.mark.parametrize("key", CASES)
def test_scalable_oracle_matches_golden(key):
    m, n, r = _case(key)
    A = oracle.uniform_matrix(m, n, int(GOLD[key + "_seed"][0]))
    P, Q = oracle.block_qr(A, r)                                 # factored-form (W, Y) evaluation
    assert np.abs(P - GOLD[key + "_packed"]).max() <= 1e-5 * np.abs(GOLD[key + "_packed"]).max()
    assert np.abs(Q - GOLD[key + "_Q"]).max() <= 1e-5
    W1, Y1 = oracle.wy_factors(GOLD[key + "_panel0"].copy(), 0, r)
    W0, Y0 = oracle.wy_transform(GOLD[key + "_panel0"].copy(), 0, r)
    assert np.array_equal(Y0, Y1) and np.abs(W0 - W1).max() <= 2e-5


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("m,n,r", REF_SHAPES[:18])
def test_oracle_vs_live_reference(m, n, r):
    A = oracle.uniform_matrix(m, n, 7 * m + n)
    P, Q = oracle.block_qr(A, r, dense=True)
    Pr, Qr = oracle.ref_block_qr(A, r)
    assert np.array_equal(P, Pr) and np.array_equal(Q, Qr)


@pytest.mark.parametrize("m,n,r", REF_SHAPES)
def test_reference_pass_criteria_hold_for_oracle(m, n, r):
    # Cuda/qr.cu:120-129, :139-158: err <= m * 2^-23 on the reference's own shape list
    A = oracle.uniform_matrix(m, n, m * n + r)
    P, Q = oracle.block_qr(A, r)
    assert oracle.backward_error(A, oracle.strip_R(P), Q) <= m * 2.0 ** -23
    assert oracle.q_error_max(Q) <= m * 2.0 ** -23
    assert abs(oracle.backward_error_packed(A, P) - oracle.backward_error(A, oracle.strip_R(P), Q)) <= 1e-6


def test_python_fixtures_vs_lapack():
    # python/test_data.py:4-57; the reference asserts allclose to np.linalg.qr (test_all.py:36-37)
    fixtures = [
        np.array([[1, 2, 3], [4, 5, 6], [7, 8, 7], [4, 2, 3], [4, 2, 2]], np.float32),
        np.array([[0, 3, 1], [0, 4, -2], [2, 1, 1]], np.float32),
        np.array([[10, 20, 30, 40, 50, 60], [32, 32, 44, 55, 66, 35], [23, 66, 74, 64, 45, 65],
                  [67, 28, 46, 26, 46, 42], [95, 95, 52, 88, 65, 11], [75, 53, 96, 47, 32, 32]], np.float32),
        np.array([[1, 2, 3], [1, 2, 3], [1, 2, 3]], np.float32),
        np.array([[1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32),
        np.array([[1, 2, 3], [0, 0, 0], [0, 0, 0]], np.float32),
    ]
    for A in fixtures:
        P, Q = oracle.block_qr(A, 2)
        R = oracle.strip_R(P)
        _, Rl = np.linalg.qr(A.astype(np.float64), mode="complete")
        assert np.allclose(np.abs(R), np.abs(Rl), atol=2e-4 * np.abs(A).max())
        assert np.allclose(Q.astype(np.float64) @ R, A, atol=2e-4 * np.abs(A).max())


def test_wide_matrix_extension():
    # m < n is undefined in the reference (Cuda/qr.cu:222-230); the oracle factors min(m,n) columns
    A = oracle.uniform_matrix(20, 50, 3)
    P, Q = oracle.block_qr(A, 8)
    R = oracle.strip_R(P)
    assert oracle.backward_error(A, R, Q) < 1e-6
    _, Rl = np.linalg.qr(A.astype(np.float64))
    assert np.allclose(np.abs(R), np.abs(Rl), atol=1e-4)


def test_tsqr_oracle_reference_case():
    # python/ca_qr.py:86-92
    np.random.seed(0)
    A = np.random.random((24, 3))
    Q, R = oracle.tsqr(A, 4)
    Ql, Rl = np.linalg.qr(A)
    assert np.allclose(Q, Ql) and np.allclose(R, Rl)
    A = np.random.random((4096, 32))
    for nb in (1, 4, 16):
        Q, R = oracle.tsqr(A, nb)
        assert np.allclose(Q @ R, A) and np.allclose(Q.T @ Q, np.eye(32))


def test_flop_models():
    assert oracle.householder_flops(2048, 2048) == pytest.approx(1.145e10, rel=1e-3)
    assert oracle.householder_flops(4096, 16384) == pytest.approx(5.04e11, rel=2e-3)
    assert oracle.householder_flops(32768, 32768) == pytest.approx(4.691e13, rel=1e-3)
    assert oracle.lib().orc_ref_flop_model(2048, 2048) == pytest.approx(2.86e10, rel=2e-3)  # Cuda/qr.cu:102-113
