"""CPU: EuRoC Jacobian text loader (mpqr_read_euroc_jacobian) against the format of the reference's
read_euroc_jacobian (Cuda/qr.cu:696-776): header "<rows> <cols>", 0-based "<row> <col> <value>" triples,
zero fill, later entries overwrite earlier ones.  The real data set is an absent LFS blob, so the
file is a synthetic block-sparse "Jacobian-like" matrix written in that format."""
import numpy as np
import pytest

import mixedprecisionblockqr_b200 as pkg


def _write(path, A, extra=()):
    m, n = A.shape
    with open(path, "w") as f:
        f.write(f"{m} {n}\n")
        for i, j in zip(*np.nonzero(A)):
            f.write(f"{i} {j} {A[i, j]:.9g}\n")
        for line in extra:
            f.write(line + "\n")


def test_loader_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    m, n = 60, 27
    A = np.zeros((m, n), np.float32)
    for rb in range(0, m, 6):               # dense 6 x 9 blocks on a sparse background
        cb = int(rng.integers(0, n - 9))
        A[rb:rb + 6, cb:cb + 9] = rng.standard_normal((6, 9)).astype(np.float32)
    p = tmp_path / "A_000000100.txt"
    _write(p, A, extra=["3 4 7.5", "3 4 -2.25"])   # duplicate entry: the last one wins
    P = pkg.read_euroc_jacobian(str(p))
    assert P.shape == (m + 1, n) and P.dtype == np.float32
    A[3, 4] = -2.25
    assert np.array_equal(P[:m], A)
    assert np.all(P[m] == 0)                # the extra row of the packed layout (Cuda/qr.cu:1866-1875)


def test_loader_errors(tmp_path):
    with pytest.raises(pkg.MpqrError):
        pkg.read_euroc_jacobian(str(tmp_path / "missing.txt"))
    bad = tmp_path / "bad.txt"
    bad.write_text("4 4\n1 9 2.0\n")
    with pytest.raises(pkg.MpqrError, match="outside"):
        pkg.read_euroc_jacobian(str(bad))
    bad.write_text("4 4\n1 2 x\n")
    with pytest.raises(pkg.MpqrError, match="malformed"):
        pkg.read_euroc_jacobian(str(bad))
