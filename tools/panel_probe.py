"""Phase profile of the register-resident panel block kernel (clock64 deltas of CTA 0, thread 0).
python tools/panel_probe.py m,pw[,B,CS,RPT,wy] ...   (0 = automatic)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

L = pkg.lib()
L.mpqr_debug_panel_probe.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_int] * 8 + [ctypes.c_void_p, ctypes.c_void_p]
names = ["pass", "reduce", "exchange", "gather+scalars"]
for spec in sys.argv[1:]:
    m, pw, fb, fcs, frpt, wy = (list(map(int, spec.split(","))) + [0, 0, 0, 0])[:6]
    n = pw
    A = torch.rand(m + 1, n, device="cuda")
    best = 1e9
    for rep in range(3):
        dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
        B = A.clone()
        torch.cuda.synchronize()
        pkg.check(L.mpqr_debug_panel_probe(B.data_ptr(), n, m, n, 0, pw, fb, fcs, frpt, wy, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    steps = max(1, d[6])
    nblk = max(1, (pw + d[9] - 1) // max(1, d[9]))
    per = {names[i]: d[i] / (steps * nblk) * 1.0 for i in range(4)}
    print(f"m={m} pw={pw} B={d[9]} CS={d[7]} RPT={d[8]} (caps max_cs,cs,rpt={d[13:16]}): per-step cycles "
          + " ".join(f"{k}={v:.0f}" for k, v in per.items())
          + f" | total/step={sum(per.values()):.0f} | load={d[4] / nblk:.0f} tail={d[5] / nblk:.0f} (sum over {nblk} blocks / nblk; steps/blk={steps})", flush=True)
