"""FP32 CUDA-core GEMMs of the FP32 driver / TSQR on their own: correctness against torch (FP64) and TFLOP/s.
python tools/sgemm_time.py"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

L = pkg.lib()
L.mpqr_debug_sgemm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)


def run(op, M, N, K, reps=5):
    if op == 0:
        X = torch.randn(K, M, device="cuda"); Z = torch.randn(K, N, device="cuda"); C = torch.zeros(M, N, device="cuda")
        ref = X.double().T @ Z.double()
    else:
        X = torch.randn(M, K, device="cuda"); Z = torch.randn(K, N, device="cuda"); C = torch.randn(M, N, device="cuda")
        ref = (C.double() - X.double() @ Z.double()) if op == 1 else X.double() @ Z.double()
    C0 = C.clone()
    pkg.check(L.mpqr_debug_sgemm(op, X.data_ptr(), X.stride(0), Z.data_ptr(), Z.stride(0), C.data_ptr(), C.stride(0), M, N, K, st))
    torch.cuda.synchronize()
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    ts = []
    for _ in range(reps):
        C.copy_(C0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.mpqr_debug_sgemm(op, X.data_ptr(), X.stride(0), Z.data_ptr(), Z.stride(0), C.data_ptr(), C.stride(0), M, N, K, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"op={op} M={M} N={N} K={K}: {t * 1e3:8.1f} us  {2.0 * M * N * K / t / 1e9:6.2f} TFLOP/s  rel err {err:.2e}", flush=True)


for (op, M, N, K) in [(0, 128, 128, 32768), (0, 128, 256, 32768), (0, 256, 256, 32768), (1, 32768, 128, 128), (1, 32768, 256, 128),
                      (2, 32768, 256, 256), (0, 4096, 4096, 4096), (1, 4096, 4096, 4096), (0, 128, 2048, 2048), (1, 2048, 2048, 128),
                      (0, 100, 132, 1000), (1, 1000, 132, 100), (2, 1028, 68, 36), (0, 64, 64, 40)]:
    run(op, M, N, K)
