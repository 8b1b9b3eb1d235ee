"""Phase times of the host-pointer drop-in (mpqr_block_qr_host) on a pinned buffer: MPQR_HOST_TRACE=1 python tools/host_trace.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = n
host = torch.rand((m + 1, n), dtype=torch.float32).pin_memory()
host[m].zero_()
src = host.clone()
for it in range(3):
    host.copy_(src)
    t0 = time.perf_counter()
    pkg.check(pkg.lib().mpqr_block_qr_host(host.data_ptr(), None, m, n, 128, pkg.MPQR_FP16), "host")
    print(f"call {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
