"""Phase profile of the panel kernel (clock64 deltas of CTA 0): python tools/panel_probe.py m pw [rows_hint ...]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

L = pkg.lib()
L.mpqr_debug_panel_probe.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p]
names = ["pass", "reduce+publish", "exchange", "gather", "scalars", "load", "store", "steps", "G", "rows", "tail(gram+T)", "CS", "NC"]
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 0
for spec in sys.argv[1:]:
    m, pw, hint, wy, fcs = (list(map(int, spec.split(","))) + [0, 0, 0])[:5]
    n = pw
    A = torch.rand(m + 1, n, device="cuda")
    for rep in range(2):
        dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
        B = A.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        pkg.check(L.mpqr_debug_panel_probe(B.data_ptr(), n, m, n, 0, pw, hint, fcs, wy, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    steps = max(1, d[7])
    per = {names[i]: d[i] / steps for i in range(5)}
    print(f"m={m} pw={pw} G={d[8]} CS={d[11]} NC={d[12]} rows/cta={d[9]} wy={wy}: per-step cycles " + " ".join(f"{k}={v:.0f}" for k, v in per.items())
          + f" caps(cs,nc16,nc8)={d[13:16]} | total/step={sum(per.values()):.0f} | load={d[5]} store={d[6]} tail={d[10]} | event(ms incl. alloc)={e0.elapsed_time(e1):.3f}", flush=True)
