"""Timing of the device metric kernels on a factorisation + explicit Q.  usage: metrics_time.py n [prec]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
m = n
st = torch.cuda.current_stream().cuda_stream
A0 = torch.zeros(m, n, device="cuda")
pkg.fill_uniform(A0.data_ptr(), n, n, 0, m, 0, n, 7, st)
A = torch.zeros(m + 1, n, device="cuda"); A[:m] = A0
Q = torch.zeros(m, m, device="cuda")
plan = pkg.BlockQR(m, n, 128, precision=prec, keep_wy=True)
plan.factor(A.data_ptr(), n, st)
plan.form_q(Q.data_ptr(), m, st)
torch.cuda.synchronize()
t0 = time.perf_counter(); be, an = pkg.backward_error(A0.data_ptr(), n, A.data_ptr(), n, Q.data_ptr(), m, m, n, st); t1 = time.perf_counter()
qe = pkg.q_error(Q.data_ptr(), m, m, st); t2 = time.perf_counter()
fl_b = m * m * n            # upper-triangular R: half of 2 m^2 n
fl_q = m * m * m            # symmetric: half of 2 m^3
print(f"metrics {m}x{n} {prec}: backward {be:.3e} in {(t1 - t0) * 1e3:.1f} ms ({fl_b / (t1 - t0) / 1e12:.1f} TFLOP/s FP64), "
      f"q_error max {qe['max_abs']:.3e} fro {qe['fro']:.3e} in {(t2 - t1) * 1e3:.1f} ms ({fl_q / (t2 - t1) / 1e12:.1f} TFLOP/s FP64)", flush=True)
