import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg
stage = sys.argv[1]
st = torch.cuda.current_stream().cuda_stream
if stage.startswith("plan"):
    _, m, n, r, kw = stage.split(",")
    m, n, r = int(m), int(n), int(r)
    A = torch.rand(m + 1, n, device="cuda")
    p = pkg.BlockQR(m, n, r, precision="fp32", keep_wy=kw == "1")
    p.factor(A.data_ptr(), n, st)
    torch.cuda.synchronize()
    print(stage, "factor ok", flush=True)
else:
    _, m, n, q = stage.split(",")
    m, n = int(m), int(n)
    A = torch.rand(m, n, device="cuda")
    Q = torch.zeros(m, n, device="cuda") if q == "1" else None
    R = torch.zeros(n, n, device="cuda")
    pkg.tsqr(A.data_ptr(), n, m, n, Q.data_ptr() if Q is not None else None, n, R.data_ptr(), n, st)
    torch.cuda.synchronize()
    print(stage, "tsqr ok", flush=True)
