"""Phase profile of the register-resident panel block kernel (clock64 deltas of CTA 0, thread 0).
python tools/panel_probe.py m,pw[,B,CS,RPT,wy] ...   (0 = automatic)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

L = pkg.lib()
L.mpqr_debug_panel_probe.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_int] * 8 + [ctypes.c_void_p, ctypes.c_void_p]
names = ["pass", "shfl", "smem+bar", "sum+send", "xchg", "scalars"]
for spec in sys.argv[1:]:
    m, pw, fb, fcs, frpt, wy = (list(map(int, spec.split(","))) + [0, 0, 0, 0])[:6]
    n = pw
    A = torch.rand(m + 1, n, device="cuda")
    try:
        for rep in range(3):
            dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
            B = A.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pkg.check(L.mpqr_debug_panel_probe(B.data_ptr(), n, m, n, 0, pw, fb, fcs, frpt, wy, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
            e1.record()
            torch.cuda.synchronize()
    except pkg.MpqrError as e:
        print(f"m={m} pw={pw} B={fb} CS={fcs} RPT={frpt}: {e}", flush=True)
        continue
    d = dbg.cpu().tolist()
    steps = max(1, d[8])
    bw = max(1, d[11])
    nblk = max(1, (pw + bw - 1) // bw)
    per = {names[i]: d[i] / (steps * nblk) for i in range(6)}
    print(f"m={m} pw={pw} B={d[11]} CS={d[9]} RPT={d[10]}: per-step cycles "
          + " ".join(f"{k}={v:.0f}" for k, v in per.items())
          + f" | total/step={sum(per.values()):.0f} | load={d[6] / nblk:.0f} tail={d[7] / nblk:.0f} per block ({nblk} blocks, {steps} steps/blk)", flush=True)
