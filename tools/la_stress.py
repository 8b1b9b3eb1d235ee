"""Stress of the look-ahead driver (green-context partitions, in-block and register-block look-ahead): every shape is
factored several times; a race between the streams would show up as a sporadic outlier of the sampled backward error
or of the |R| difference between repetitions.  usage: MPQR_OVERLAP=1 python tools/la_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg
from tools.quick_time import sampled_backward_error

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
st = torch.cuda.current_stream().cuda_stream
shapes = [(16384, 16384, 128, 0), (8192, 8192, 128, 0), (12000, 9000, 128, 1024), (8192, 8192, 64, 512), (6000, 16000, 96, 768),
          (4096, 4096, 128, 512), (9000, 5000, 128, 640), (20000, 6000, 128, 1024), (5000, 5000, 32, 512)]
bad = 0
for (m, n, r, nb) in shapes:
    lda = (n + 7) // 8 * 8
    A0 = torch.zeros(m, lda, device="cuda")
    pkg.fill_uniform(A0.data_ptr(), lda, n, 0, m, 0, n, 7 * m + n, st)
    A = torch.zeros(m + 1, lda, device="cuda")
    plan = pkg.BlockQR(m, n, r, nb=nb, precision="fp16")
    bes, Rs = [], []
    for it in range(reps):
        A[:m].copy_(A0); A[m].zero_()
        plan.factor(A.data_ptr(), lda, st)
        torch.cuda.synchronize()
        bes.append(sampled_backward_error(A0[:, :n], A[:, :n], r=plan.r))
        Rs.append(torch.triu(A[:min(m, n), :n]).abs().clone())
    dr = max(((Rs[i] - Rs[0]).abs().max() / Rs[0].max()).item() for i in range(1, reps)) if reps > 1 else 0.0
    ok = max(bes) <= 1.3 * min(bes) and max(bes) < 12 * 2.0 ** -11 and dr < 60 * 2.0 ** -11
    bad += not ok
    print(f"{m}x{n} r={plan.r} nb={plan.nb}: backward {min(bes):.2e}..{max(bes):.2e}  |R| spread between reps {dr:.1e}  launches={plan.last_launches}  {'ok' if ok else 'OUTLIER'}", flush=True)
    plan.close()
    del A, A0
    torch.cuda.empty_cache()
print("stress:", "all ok" if bad == 0 else f"{bad} shapes with outliers")
