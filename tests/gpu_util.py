"""Helpers for the -m gpu tests: device buffers via torch (plumbing only), calls via the C-ABI."""
import numpy as np
import torch

import mixedprecisionblockqr_b200 as pkg


def stream():
    return torch.cuda.current_stream().cuda_stream


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def panel_factor(P, lam, pw, want_wy=True):
    """P: packed (m+1) x n numpy.  Returns (packed, Y, W, T) numpy."""
    m, n = P.shape[0] - 1, P.shape[1]
    dA = to_dev(P)
    D = m - lam
    dY = torch.zeros(D, pw, device="cuda") if want_wy else None
    dW = torch.zeros(D, pw, device="cuda") if want_wy else None
    dT = torch.zeros(pw, pw, device="cuda") if want_wy else None
    rc = pkg.lib().mpqr_panel_factor_device(dA.data_ptr(), n, m, n, lam, pw,
                                            dY.data_ptr() if want_wy else None,
                                            dW.data_ptr() if want_wy else None,
                                            dT.data_ptr() if want_wy else None, stream())
    pkg.check(rc, "mpqr_panel_factor_device")
    torch.cuda.synchronize()
    if want_wy:
        return dA.cpu().numpy(), dY.cpu().numpy(), dW.cpu().numpy(), dT.cpu().numpy()
    return dA.cpu().numpy(), None, None, None


def gemm_tn(X, Z, bf16=False, xoff=0, zoff=0):
    """X [K x M], Z [K x N] float arrays with values exactly representable in 16 bit."""
    dt = torch.bfloat16 if bf16 else torch.float16
    K, M = X.shape
    N = Z.shape[1]
    ldx, ldz = ((M + xoff + 7) // 8) * 8, ((N + zoff + 7) // 8) * 8
    dX = torch.zeros(K, ldx, dtype=dt, device="cuda")
    dZ = torch.zeros(K, ldz, dtype=dt, device="cuda")
    dX[:, xoff:xoff + M] = torch.from_numpy(X).to(dt)
    dZ[:, zoff:zoff + N] = torch.from_numpy(Z).to(dt)
    lds = ((N + 3) // 4) * 4
    dS = torch.full((M, lds), 7.0, device="cuda")
    rc = pkg.lib().mpqr_gemm_tn_device(dX.data_ptr() + 2 * xoff, ldx, dZ.data_ptr() + 2 * zoff, ldz,
                                       dS.data_ptr(), lds, M, N, K, int(bf16), stream())
    pkg.check(rc, "mpqr_gemm_tn_device")
    torch.cuda.synchronize()
    return dS.cpu().numpy()[:, :N], dS.cpu().numpy()[:, N:]


def gemm_nn(X, S, C, bf16=False, shadow=True, coff=0):
    dt = torch.bfloat16 if bf16 else torch.float16
    M, K = X.shape
    N = S.shape[1]
    ldx, lds = ((K + 7) // 8) * 8, ((N + 7) // 8) * 8
    ldc = ((N + coff + 7) // 8) * 8
    dX = torch.zeros(M, ldx, dtype=dt, device="cuda")
    dS = torch.zeros(K, lds, dtype=dt, device="cuda")
    dX[:, :K] = torch.from_numpy(X).to(dt)
    dS[:, :N] = torch.from_numpy(S).to(dt)
    dC = torch.full((M, ldc), 3.0, device="cuda")
    dC[:, coff:coff + N] = torch.from_numpy(C).cuda()
    dH = torch.full((M, ldc), 5.0, dtype=dt, device="cuda")
    rc = pkg.lib().mpqr_gemm_nn_device(dX.data_ptr(), ldx, dS.data_ptr(), lds, dC.data_ptr() + 4 * coff, ldc,
                                       (dH.data_ptr() + 2 * coff) if shadow else None, ldc, M, N, K, int(bf16), stream())
    pkg.check(rc, "mpqr_gemm_nn_device")
    torch.cuda.synchronize()
    return dC.cpu().numpy(), dH.float().cpu().numpy()


# ------------------------------------------------------------------ whole-factorisation helpers (GPU side, FP64)
def sampled_backward_error(A0, P, r, k=16):
    """||(A - QR) X||_F / (||A||_F sqrt(k)) with Gaussian X, FP64 on the device, straight from the packed factor
    (torch tensors: A0 m x n, P (m+1) x n).  Q is applied panel by panel from the stored unit vectors, so this is the
    reference's ||A - QR||_F / ||A||_F (Cuda/qr.cu:115-135) sampled on k random directions."""
    m, n = A0.shape
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, k, device="cuda", dtype=torch.float64, generator=g)
    AX = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    Z = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    an2 = 0.0
    step = 4096
    for i in range(0, m, step):
        blk = A0[i:i + step].double()
        AX[i:i + step] = blk @ X
        an2 += float((blk * blk).sum())
        Z[i:i + step] = torch.triu(P[i:min(i + step, m)].double(), diagonal=i) @ X
    kmax = min(m, n)
    for lam in range(((kmax - 1) // r) * r, -1, -r):
        pw = min(r, kmax - lam)
        Y = torch.tril(P[lam + 1:m + 1, lam:lam + pw].double())
        Tinv = torch.triu(Y.T @ Y, 1) + 0.5 * torch.eye(pw, device="cuda", dtype=torch.float64)
        Z[lam:] -= Y @ torch.linalg.solve_triangular(Tinv, Y.T @ Z[lam:], upper=True)
    return float(torch.linalg.norm(AX - Z)) / (an2 ** 0.5 * k ** 0.5)


def factor_device(A, r, precision="fp16", nb=0, reps=1):
    """Factors the numpy matrix A through the plan API with device-resident buffers; returns (A0 device m x n view,
    packed factor as a device tensor (m+1) x n view, plan.r, plan.nb)."""
    m, n = A.shape
    lda = (n + 7) // 8 * 8
    A0 = torch.zeros(m, lda, device="cuda")
    A0[:, :n] = torch.from_numpy(np.ascontiguousarray(A, np.float32)).cuda()
    dA = torch.zeros(m + 1, lda, device="cuda")
    plan = pkg.BlockQR(m, n, r, nb=nb, precision=precision)
    for _ in range(reps):
        dA[:m].copy_(A0)
        dA[m].zero_()
        plan.factor(dA.data_ptr(), lda, stream())
        torch.cuda.synchronize()
    rr, nbb = plan.r, plan.nb
    plan.close()
    return A0[:, :n], dA[:, :n], rr, nbb


def r_rel_diff(P, Rref):
    """max | |R| - |R_ref| | / max |R_ref| with R taken from the packed factor (device tensor or numpy)."""
    Pn = P.cpu().numpy() if hasattr(P, "cpu") else P
    m, n = Rref.shape[0], Rref.shape[1]
    R = np.triu(Pn[:m, :n])[:Rref.shape[0]]
    return float(np.abs(np.abs(R) - np.abs(Rref)).max() / np.abs(Rref).max())


def observe(name, **vals):
    """Appends the observed figures of a test to gpurun_out/observed.jsonl (tolerances are kept at <= 3x of these)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "observed.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: float(v) for k, v in vals.items()}}) + "\n")
