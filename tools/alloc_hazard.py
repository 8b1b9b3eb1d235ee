"""Does a device allocation issued while a factorisation is in flight stall the persistent panel kernel for good?
(include/mpqr.h, MPQR_STREAM_ORDERED).  usage: alloc_hazard.py [ordered]   -- run under tools/runs/watch.sh"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

ordered = len(sys.argv) > 1 and sys.argv[1] == "ordered"
m = n = 16384
st = torch.cuda.current_stream().cuda_stream
plan = pkg.BlockQR(m, n, 128, stream_ordered=ordered)
A0 = torch.zeros(m, n, device="cuda")
pkg.fill_uniform(A0.data_ptr(), n, n, 0, m, 0, n, 7, st)
A = torch.zeros(m + 1, n, device="cuda")
rt = ctypes.CDLL("libcudart.so.12") if False else None
drv = ctypes.CDLL("libcuda.so.1")
for w in range(2):
    A[:m].copy_(A0); plan.factor(A.data_ptr(), n, st)
torch.cuda.synchronize()
print("warm-up done, ordered =", ordered, flush=True)
for trial in range(4):
    A[:m].copy_(A0)
    t0 = time.perf_counter()
    plan.factor(A.data_ptr(), n, st)
    time.sleep(0.003 * (trial + 1))          # the chain is in flight now
    ptrs = []
    for k in range(4):                        # fresh device allocations (driver API: not served from torch's cache)
        p = ctypes.c_uint64()
        rc = drv.cuMemAlloc_v2(ctypes.byref(p), ctypes.c_size_t((256 + 64 * k) << 20))
        ptrs.append(p)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for p in ptrs:
        drv.cuMemFree_v2(p)
    print(f"trial {trial}: issue+alloc {1e3 * (t1 - t0):.1f} ms, drained after {1e3 * (t2 - t0):.1f} ms", flush=True)
print("no hang", flush=True)
