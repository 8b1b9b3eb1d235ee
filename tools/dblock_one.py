import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg
st = torch.cuda.current_stream().cuda_stream
m, pw, lda = 32768, 32, 32
A = torch.rand(m + 1, lda, device="cuda")
plan = pkg.BlockQR(m, pw, pw, precision="fp32")
for _ in range(3):
    plan.factor(A.data_ptr(), lda, st)
torch.cuda.synchronize()
print("ok")
