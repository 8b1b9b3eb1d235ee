"""Generates tests/golden/ref_block_qr.npz by running the UNMODIFIED reference (oracle/_ref/libref_qr.so,
built by oracle/build_ref.sh from /root/reference/Cuda/{qr.cu,mmult.cu}) on seeded inputs.
Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Stored per case: the input seed/shape, the packed factor (m+1 x n) and Q (m x m) of reference
h_block_qr (Cuda/qr.cu:1275), the panel factor of h_householder_qr (Cuda/qr.cu:198) for the first
panel, the dense I - W Y^T of h_wy_transform (Cuda/qr.cu:337), and the explicit Q of
h_q_backward_accumulation (Cuda/qr.cu:296)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [(6, 4, 2), (12, 8, 5), (12, 8, 8), (24, 16, 12), (60, 40, 16), (97, 90, 16), (129, 80, 16), (80, 80, 16)]

out = {}
for (m, n, r) in CASES:
    seed = 1000 * m + n + r
    A = oracle.uniform_matrix(m, n, seed)
    P, Q = oracle.ref_block_qr(A, r)
    key = f"{m}x{n}r{r}"
    out[key + "_seed"] = np.array([seed])
    out[key + "_packed"] = P
    out[key + "_Q"] = Q
    P1 = oracle.ref_householder_panel(oracle.pack(A), 0, r)
    out[key + "_panel0"] = P1
    out[key + "_wydense0"] = oracle.ref_wy_dense(P1.copy(), 0, r)
    Pfull = oracle.ref_householder_panel(oracle.pack(A), 0, n)
    out[key + "_hhfull"] = Pfull
    out[key + "_qback"] = oracle.ref_q_backward_accumulation(Pfull.copy())
# reference known answer, Cuda/qr.cu:1397-1401 (3x3), produced by the reference itself with r=2
A3 = np.array([[12, -51, 4], [6, 167, -68], [-4, 24, -41]], np.float32)
P3, Q3 = oracle.ref_block_qr(A3, 2)
out["known3x3_packed"], out["known3x3_Q"] = P3, Q3
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_block_qr.npz"), **out)
print("wrote", len(out), "arrays")
