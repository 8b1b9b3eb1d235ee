"""Wall time of mpqr_tsqr_device (allocation + lanes + factor + thin Q, device-synchronous call) for a
tall-skinny matrix, and of the device metric kernels.  usage: tsqr_time.py [m n] ; MPQR_TSQR_LANES=k"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1048576, 256)
st = torch.cuda.current_stream().cuda_stream
A = torch.zeros(m, n, device="cuda")
pkg.fill_uniform(A.data_ptr(), n, n, 0, m, 0, n, 1048576256, st)
Q = torch.zeros(m, n, device="cuda")
R = torch.zeros(n, n, device="cuda")
F = pkg.householder_flops(m, n)
for want_q in (True, False):
    ts = []
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pkg.tsqr(A.data_ptr(), n, m, n, Q.data_ptr() if want_q else None, n, R.data_ptr(), n, st)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    best = min(ts)
    gb = (4.0 * m * n * (2 if want_q else 1)) / 1e9
    print(f"tsqr {m}x{n} lanes={os.environ.get('MPQR_TSQR_LANES', '8')} Q={want_q}: {best:.2f} ms  {F / best / 1e9:.2f} TFLOP/s  "
          f"{gb / best * 1e3:.0f} GB/s algorithmic  all={['%.1f' % t for t in ts]}", flush=True)
be = (Q.double().T @ Q.double() - torch.eye(n, device="cuda", dtype=torch.float64)).norm().item()
res = ((A.double() - Q.double() @ R.double()).norm() / A.double().norm()).item()
print(f"  orth {be:.2e}  backward {res:.2e}", flush=True)
