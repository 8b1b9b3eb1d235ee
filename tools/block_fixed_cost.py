"""Fixed cost of one register-block kernel launch: FP32 plan on an m x pw matrix (one launch per factor call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg
st = torch.cuda.current_stream().cuda_stream
for m in (32768, 16384, 4096):
    for pw in (1, 2, 4, 8, 16, 32):
        lda = 32
        A0 = torch.rand(m + 1, lda, device="cuda")
        A = A0.clone()
        plan = pkg.BlockQR(m, pw, pw, precision="fp32")
        for _ in range(5):
            plan.factor(A.data_ptr(), lda, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            plan.factor(A.data_ptr(), lda, st)
        e1.record()
        torch.cuda.synchronize()
        print(f"m={m} pw={pw}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per factor call ({plan.last_launches} launches)", flush=True)
        plan.close()
