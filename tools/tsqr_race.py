"""Repeats mpqr_tsqr_device on a few shapes and reports the |R| deviation from an FP64 QR (GPU, torch) each time:
a race between the lanes shows up as sporadic large deviations.  usage: tsqr_race.py reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
st = torch.cuda.current_stream().cuda_stream
for (m, n) in [(65836, 128), (100000, 256), (70000, 32), (300000, 64)]:
    A = torch.zeros(m, n, device="cuda")
    pkg.fill_uniform(A.data_ptr(), n, n, 0, m, 0, n, m + n, st)
    Rl = torch.linalg.qr(A.double(), mode="r").R.abs()
    out = []
    for it in range(reps):
        Q = torch.zeros(m, n, device="cuda")
        R = torch.zeros(n, n, device="cuda")
        pkg.tsqr(A.data_ptr(), n, m, n, Q.data_ptr(), n, R.data_ptr(), n, st)
        torch.cuda.synchronize()
        d = ((R.double().abs() - Rl).abs().max() / Rl.max()).item()
        be = ((A.double() - Q.double() @ R.double()).norm() / A.double().norm()).item()
        out.append(f"{d:.1e}/{be:.1e}")
    print(f"{m}x{n} lanes={os.environ.get('MPQR_TSQR_LANES', '8')} pdl={'off' if os.environ.get('MPQR_NO_PDL') else 'on'}: |R| dev / backward: {' '.join(out)}", flush=True)
