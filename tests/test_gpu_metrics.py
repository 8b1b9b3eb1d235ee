"""Device metric kernels (mpqr_*_error_device) against the oracle's FP64 restatement
of the reference's harness metrics (oracle/mpqr_oracle.c: orc_backward_error / orc_q_error_max /
orc_orthogonality_fro / orc_strip_R, following Cuda/qr.cu:85-196) and, where the reference itself
compiled (oracle/_ref), against its own FP32 host functions within their FP32 rounding."""
import ctypes

import numpy as np
import pytest
import torch

import mixedprecisionblockqr_b200 as pkg
import oracle

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


def _factor(m, n, r, seed):
    """Oracle factorisation (packed factor + explicit Q) of a seeded uniform matrix."""
    A = oracle.uniform_matrix(m, n, seed)
    P, Q = oracle.block_qr(A, r, want_q=True)
    return A, P, Q


@pytest.mark.parametrize("m,n,r", [(3, 3, 1), (60, 40, 8), (97, 90, 16), (129, 80, 16), (200, 200, 32), (130, 257, 16)])
def test_metrics_vs_oracle(m, n, r):
    if m < n:
        # the reference drivers do not support m < n; use a square factorisation of the left part
        A, P, Q = _factor(m, m, r, 11 * m + n)
        extra = oracle.uniform_matrix(m, n - m, 5)
        A = np.concatenate([A, extra], axis=1)
        P = np.concatenate([P, np.concatenate([Q.T @ extra, np.zeros((1, n - m), np.float32)])], axis=1).astype(np.float32)
        P = np.ascontiguousarray(P)
    else:
        A, P, Q = _factor(m, n, r, 11 * m + n)
    R = oracle.strip_R(P)
    dA0 = torch.from_numpy(A).cuda()
    dP = torch.from_numpy(P).cuda()
    dQ = torch.from_numpy(Q).cuda()
    dR = torch.full((m, n), 7.0, device="cuda")
    pkg.strip_R(dP.data_ptr(), n, dR.data_ptr(), n, m, n, _st())
    torch.cuda.synchronize()
    assert np.array_equal(dR.cpu().numpy(), R)                      # bit-exact copy / mask
    # the packed factor may be passed where R is expected (masked read)
    be_packed, an = pkg.backward_error(dA0.data_ptr(), n, dP.data_ptr(), n, dQ.data_ptr(), m, m, n, _st())
    be_r, _ = pkg.backward_error(dA0.data_ptr(), n, dR.data_ptr(), n, dQ.data_ptr(), m, m, n, _st())
    be_o = oracle.backward_error(A, R, Q)
    assert be_packed == be_r
    assert abs(be_packed - be_o) <= 1e-9 * max(be_o, 1e-30) + 1e-15
    assert abs(an - np.linalg.norm(A.astype(np.float64))) <= 1e-12 * an
    qe = pkg.q_error(dQ.data_ptr(), m, m, _st())
    assert abs(qe["max_signed"] - oracle.q_error_max(Q)) <= 1e-12
    assert abs(qe["fro"] - oracle.orthogonality_fro(Q)) <= 1e-9 * qe["fro"] + 1e-15
    G = Q.astype(np.float64).T @ Q.astype(np.float64) - np.eye(m)
    assert abs(qe["max_abs"] - np.abs(G).max()) <= 1e-12
    # ||strict lower part||: 0 on the stripped R, the Householder vectors' norm on the packed factor
    assert pkg.lower_trapezoid_error(dR.data_ptr(), n, m, n, _st()) == 0.0
    lt = pkg.lower_trapezoid_error(dP.data_ptr(), n, m, n, _st())
    assert abs(lt - np.linalg.norm(np.tril(P[:m].astype(np.float64), -1))) <= 1e-12 * max(lt, 1.0)
    fn = pkg.frobenius_norm(dP.data_ptr(), n, m + 1, n, _st())
    assert abs(fn - np.linalg.norm(P.astype(np.float64))) <= 1e-12 * fn


def test_metrics_strided_and_perturbed():
    """Leading dimensions larger than the widths; a perturbed R shows up at the right size."""
    m, n, r = 150, 70, 16
    A, P, Q = _factor(m, n, r, 99)
    R = oracle.strip_R(P)
    R2 = R.copy()
    R2[3, 40] += 0.25
    lda, ldr, ldq = n + 5, n + 9, m + 3
    dA0 = torch.zeros(m, lda, device="cuda"); dA0[:, :n] = torch.from_numpy(A).cuda()
    dR = torch.full((m, ldr), 3.0, device="cuda"); dR[:, :n] = torch.from_numpy(R2).cuda()
    dQ = torch.full((m, ldq), 3.0, device="cuda"); dQ[:, :m] = torch.from_numpy(Q).cuda()
    be, _ = pkg.backward_error(dA0.data_ptr(), lda, dR.data_ptr(), ldr, dQ.data_ptr(), ldq, m, n, _st())
    assert abs(be - oracle.backward_error(A, R2, Q)) <= 1e-9 * be
    assert be > 0.2 / np.linalg.norm(A)
    dRr = torch.from_numpy(R).cuda()
    ag = pkg.r_agreement(dR.data_ptr(), ldr, dRr.data_ptr(), n, m, n, _st())
    assert abs(ag["max_abs_diff"] - 0.25) < 1e-6 and abs(ag["fro_diff"] - 0.25) < 1e-6
    assert abs(ag["max_abs_ref"] - np.abs(R).max()) == 0
    # sign-insensitive: flipping a row of R (the other Householder sign convention) changes nothing
    dRf = dRr.clone(); dRf[5] *= -1
    ag = pkg.r_agreement(dRf.data_ptr(), n, dRr.data_ptr(), n, m, n, _st())
    assert ag["max_abs_diff"] == 0.0


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_metrics_vs_reference_host_functions():
    """The reference's own FP32 host metrics (Cuda/qr.cu:115-196) on the same Q, R: equal within FP32 rounding."""
    m, n, r = 96, 64, 16
    A, P, Q = _factor(m, n, r, 4242)
    R = oracle.strip_R(P)
    ref = oracle.ref()
    f = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    be_ref = ref.ref_h_backward_error(f(A), f(R), f(Q), m, n, 23)
    qe_ref = ref.ref_h_q_error(f(Q), m, 23)
    dA0, dR, dQ = (torch.from_numpy(x).cuda() for x in (A, R, Q))
    be, _ = pkg.backward_error(dA0.data_ptr(), n, dR.data_ptr(), n, dQ.data_ptr(), m, m, n, _st())
    qe = pkg.q_error(dQ.data_ptr(), m, m, _st())
    assert abs(be - be_ref) <= 2e-7          # both ~1e-7: FP32 evaluation noise of the reference's metric
    assert abs(qe["max_signed"] - qe_ref) <= 5e-7


def test_metrics_on_device_factorisation_2048():
    """End to end at C2: factor + explicit Q on the GPU, metrics on the GPU; the reference's pass bounds
    m * 2^-11 (mixed) / m * 2^-23 (FP32) (Cuda/qr.cu:120-129, :1836, :1889) and far tighter in practice."""
    m = n = 2048
    lda = n
    for prec, r, bound in (("fp16", 32, 5e-3), ("fp32", 32, 5e-6)):
        A0 = torch.zeros(m, lda, device="cuda")
        pkg.fill_uniform(A0.data_ptr(), lda, n, 0, m, 0, n, 2048032, _st())
        A = torch.zeros(m + 1, lda, device="cuda"); A[:m] = A0
        Q = torch.zeros(m, m, device="cuda")
        plan = pkg.BlockQR(m, n, r, precision=prec, keep_wy=True)
        plan.factor(A.data_ptr(), lda, _st())
        plan.form_q(Q.data_ptr(), m, _st())
        be, _ = pkg.backward_error(A0.data_ptr(), lda, A.data_ptr(), lda, Q.data_ptr(), m, m, n, _st())
        qe = pkg.q_error(Q.data_ptr(), m, m, _st())
        plan.close()
        assert be <= bound, (prec, be)
        assert qe["max_abs"] <= bound * 4, (prec, qe)
        assert qe["max_signed"] <= m * 2.0 ** (-11 if prec == "fp16" else -23)
