"""Per-block timeline of the persistent panel chain (globaltimer stamps of CTA 0 / thread 0):
python tools/chain_probe.py m,pw[,cs,rpt] ...   -> us per phase: wait(far flag) load near steps T out fence+post"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

L = pkg.lib()
L.mpqr_debug_panel_probe.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_int] * 8 + [ctypes.c_void_p, ctypes.c_void_p]
names = ["wait", "load", "near", "steps", "T", "out", "post"]
for spec in sys.argv[1:]:
    vals = list(map(int, spec.split(",")))
    m, pw = vals[:2]
    fcs, frpt = (vals[2] if len(vals) > 2 else 0), (vals[3] if len(vals) > 3 else 0)
    n = pw
    A = torch.rand(m + 1, n, device="cuda")
    for rep in range(3):
        dbg = torch.zeros(8 * 8, dtype=torch.int64, device="cuda")
        B = A.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg.check(L.mpqr_debug_panel_probe(B.data_ptr(), n, m, n, 0, pw, -1, fcs, frpt, 1, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
    d = dbg.cpu().view(8, 8).tolist()
    nblk = pw // 16
    print(f"m={m} pw={pw} cs={fcs} rpt={frpt}: whole call {e0.elapsed_time(e1) * 1e3:.0f} us (chain kernel + side updates + finalize + Gram/T/W); kernel span "
          f"{(d[nblk - 1][7] - d[0][0]) / 1e3:.1f} us", flush=True)
    for jb in range(nblk):
        ph = [(d[jb][k + 1] - d[jb][k]) / 1e3 for k in range(7)]
        print(f"   block {jb}: " + " ".join(f"{nm}={v:5.1f}" for nm, v in zip(names, ph)) + f" | total {(d[jb][7] - d[jb][0]) / 1e3:5.1f} us", flush=True)
