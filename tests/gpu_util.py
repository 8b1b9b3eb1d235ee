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
