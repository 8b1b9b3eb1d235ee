"""2-GPU column-block-cyclic parity (skipped when fewer than 2 GPUs are visible)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("m,n,r,nb", [(2048, 2048, 64, 256), (3000, 1600, 32, 128), (1024, 4096, 64, 256),
                                      (1000, 4096, 64, 256),     # m % nb != 0, m < n: the owner's columns right of the last reflector
                                      (6144, 6144, 128, 512)])   # 12 outer blocks: green-context partitions + persistent panel chain
def test_two_gpu_matches_single(m, n, r, nb):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mg_worker.py"), str(m), str(n), str(r), str(nb)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("m,n", [(20000, 64), (150000, 256)])
def test_two_gpu_tsqr(m, n):
    """Row-block TSQR over 2 GPUs (mpqr_mg_tsqr_device, one ncclAllGather of the R factors) against LAPACK."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "mg_tsqr_worker.py"), str(m), str(n)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
