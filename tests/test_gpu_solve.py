"""Least-squares solve on the factorisation (mpqr_solve_device; the reference's dev_QR_Solver,
Cuda/QR/Solver/solver.cu:39-87, is a stub; python/linear_least_sqare.py:5-22 is its NumPy demo):
x = R^-1 Q^T b against numpy.linalg.lstsq in FP64."""
import numpy as np
import pytest
import torch

import mixedprecisionblockqr_b200 as pkg
import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,r,nrhs,prec,tol", [
    (300, 120, 16, 1, "fp32", 2e-4), (1500, 700, 64, 3, "fp32", 2e-4), (2048, 2048, 32, 2, "fp32", 5e-3),
    (1500, 700, 64, 8, "fp16", 2e-2), (5000, 300, 128, 4, "fp32", 2e-4),
])
def test_solve_vs_lstsq(m, n, r, nrhs, prec, tol):
    A = oracle.uniform_matrix(m, n, 5 * m + n)
    rng = np.random.default_rng(m + n)
    B = rng.standard_normal((m, nrhs)).astype(np.float32)
    xref = np.linalg.lstsq(A.astype(np.float64), B.astype(np.float64), rcond=None)[0]
    lda = (n + 7) // 8 * 8
    dA = torch.zeros(m + 1, lda, device="cuda")
    dA[:m, :n] = torch.from_numpy(A).cuda()
    dB = torch.from_numpy(B).cuda()
    st = torch.cuda.current_stream().cuda_stream
    plan = pkg.BlockQR(m, n, r, precision=prec)
    plan.factor(dA.data_ptr(), lda, st)
    plan.solve(dA.data_ptr(), lda, dB.data_ptr(), nrhs, nrhs, st)
    torch.cuda.synchronize()
    out = dB.cpu().numpy().astype(np.float64)
    x = out[:n]
    assert np.linalg.norm(x - xref) <= tol * np.linalg.norm(xref)
    # the tail of Q^T b carries the residual norm
    res = np.linalg.norm(A.astype(np.float64) @ xref - B.astype(np.float64), axis=0)
    assert np.allclose(np.linalg.norm(out[n:], axis=0), res, rtol=max(tol, 1e-3), atol=1e-3)
    plan.close()
