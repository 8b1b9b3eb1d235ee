"""CPU: the C-ABI library loads without a GPU, exports every symbol include/mpqr.h declares, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import os
import re

import numpy as np
import pytest

import mixedprecisionblockqr_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "mpqr.h")).read()
    declared = set(re.findall(r"\b(mpqr_[A-Za-z0-9_]+)\s*\(", hdr)) - {"mpqr_handle"}
    assert declared == set(pkg.ABI_SYMBOLS), declared ^ set(pkg.ABI_SYMBOLS)
    L = pkg.lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in L.mpqr_version()


def test_header_flags_match_the_python_mirror():
    # flag and error constants of include/mpqr.h against the ctypes mirror (a drifted bit would silently change a plan)
    hdr = open(os.path.join(ROOT, "include", "mpqr.h")).read()
    defs = {k: int(v.rstrip("u"), 0) for k, v in re.findall(r"#define\s+(MPQR_[A-Z0-9_]+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+)u?)\)?", hdr)}
    for name in ("MPQR_FP32", "MPQR_FP16", "MPQR_BF16", "MPQR_KEEP_WY", "MPQR_STREAM_ORDERED"):
        assert defs[name] == getattr(pkg, name), name
    assert defs["MPQR_NCCL_UID_BYTES"] == pkg.NCCL_UID_BYTES
    flags = [defs[n] for n in ("MPQR_FP16", "MPQR_BF16", "MPQR_KEEP_WY", "MPQR_STREAM_ORDERED")]
    assert len({f for f in flags}) == len(flags) and all(f & (f - 1) == 0 for f in flags)   # distinct single bits
    assert defs["MPQR_PRECISION_MASK"] == defs["MPQR_FP16"] | defs["MPQR_BF16"]


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    A = np.zeros((5, 4), np.float32)
    with pytest.raises(pkg.MpqrError):
        pkg.dev_mixed_precision_block_qr(A, None, 4, 4, 2)
    with pytest.raises(pkg.MpqrError):
        pkg.BlockQR(64, 64, 16)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mixedprecisionblockqr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "oracle/" not in src.replace("oracle/mpqr_oracle.c:orc_uniform01", ""), f


def test_argument_validation_messages():
    L = pkg.lib()
    assert L.mpqr_block_qr_host(None, None, 4, 4, 2, 0) == -1
    assert b"bad arguments" in L.mpqr_last_error()


def test_flop_formula():
    assert pkg.householder_flops(32768, 32768) == pytest.approx(4.691e13, rel=1e-3)
    assert pkg.householder_flops(4096, 16384) > 0


REF_SYMBOLS = {  # C++-mangled names of the reference drivers, Cuda/qr.cuh:129-137
    "dev_mixed_precision_block_qr": "_Z28dev_mixed_precision_block_qrPfS_iii",
    "dev_block_qr_wy": "_Z15dev_block_qr_wyPfS_iii",
    "dev_block_qr": "_Z12dev_block_qrPfS_iii",
}


def test_reference_symbol_shim_exports():
    """libmpqr_refshim.so defines the reference's own driver symbols (drop-in at link level)."""
    import ctypes
    shim = os.path.join(ROOT, "mixedprecisionblockqr_b200", "libmpqr_refshim.so")
    assert os.path.exists(shim), "run python -m mixedprecisionblockqr_b200.build"
    L = ctypes.CDLL(shim)
    for sym in REF_SYMBOLS.values():
        assert hasattr(L, sym), sym


def test_reference_log_format_and_flop_model(tmp_path):
    """mpqr_write_results_to_log / mpqr_qr_flops_per_second (host only): same file format and the same
    operation count as h_write_results_to_log / h_qr_flops_per_second (reference Cuda/qr.cu:58-83, :102-113);
    the file parses the way the reference's Cuda/performance/util.py:19-31 reads it."""
    import csv
    import mixedprecisionblockqr_b200 as pkg
    d = str(tmp_path / "log")
    pkg.write_results_to_log(2048, 1024, 12.5, 3.0e9, 1.25e-4, file_name="dev_mixed", log_dir=d)
    pkg.write_results_to_log(4096, 4096, 100.0, 5.0e10, 2.0e-4, file_name="dev_mixed", log_dir=d)
    rows = list(csv.reader(open(tmp_path / "log" / "dev_mixed.txt", newline="")))
    assert rows[0] == ["rows", "cols", "runtime", "flops", "error"]
    assert len(rows) == 3 and all(len(r) == 5 for r in rows)
    assert rows[1] == ["2048.000000", "1024.000000", "12.500000", "3000000000.000000", "0.000125"]  # std::to_string(double)
    assert int(float(rows[2][0])) == 4096 and abs(float(rows[2][4]) - 2.0e-4) < 1e-6
    for (t, m, n) in [(10.0, 2048, 2048), (3.5, 97, 90), (1000.0, 1024, 512)]:
        got = pkg.qr_flops_per_second(t, m, n)
        want = (4.0 * m * m * n - m * n * n + n ** 3 / 3.0) / (t / 1000.0)
        assert abs(got - want) <= 1e-5 * want
    import oracle
    if oracle.ref_available():
        assert pkg.qr_flops_per_second(7.0, 640, 480) == oracle.ref().ref_h_qr_flops_per_second(7.0, 640, 480)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the reference's own CPU block QR from oracle/_ref, or the oracle port) must print
    ONE JSON line with the contract's keys; it needs no GPU.  (~15 s: four 640 x 640 factorisations on one core.)"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("block-QR TFLOP/s") and d["unit"] == "TFLOP/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["backward_error"] < 640 * 2.0 ** -23          # the reference's own FP32 pass bound (Cuda/qr.cu:120-129)


def test_release_cache_without_plans_is_a_noop():
    """mpqr_release_cache / mpqr_tsqr_release_cache free the plans kept between calls; with nothing cached (and no GPU)
    they must simply succeed."""
    import mixedprecisionblockqr_b200 as pkg
    assert pkg.lib().mpqr_release_cache() == 0
    assert pkg.lib().mpqr_tsqr_release_cache() == 0
