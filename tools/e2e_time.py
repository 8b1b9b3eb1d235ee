"""End-to-end time of the host drop-in (pinned host buffers, H2D + factor + D2H inside): python tools/e2e_time.py [m n r]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

m, n, r = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (32768, 32768, 128)))
st = torch.cuda.current_stream().cuda_stream
A0 = torch.zeros(m, n, device="cuda")
pkg.fill_uniform(A0.data_ptr(), n, n, 0, m, 0, n, 32768128, st)
src = torch.zeros((m + 1, n), dtype=torch.float32).pin_memory()
src[:m].copy_(A0)
host = torch.empty_like(src).pin_memory()
F = pkg.householder_flops(m, n)
for it in range(4):
    host.copy_(src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pkg.check(pkg.lib().mpqr_block_qr_host(host.data_ptr(), None, m, n, r, pkg.MPQR_FP16), "mpqr_block_qr_host")
    dt = time.perf_counter() - t0
    print(f"call {it}: {dt * 1e3:.1f} ms  {F / dt / 1e12:.1f} TFLOP/s end to end", flush=True)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import sampled_backward_error
print("sampled backward error:", sampled_backward_error(A0, host.cuda(), 128))
