"""Device timing of the two trailing-update GEMMs at far-update shapes (CUDA events, best of reps):
python tools/gemm_time.py [M_far N_far K_blk]   default 31744 31744 1024 (first far update of 32768^2, nb = 1024)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg

D, N, KB = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (31744, 31744, 1024)))
L = pkg.lib()
st = torch.cuda.current_stream().cuda_stream
W = (torch.rand(D, KB, device="cuda") - 0.5).half()        # W / Y block: D x kb (TN: K = D rows, M = kb)
A16 = (torch.rand(D, N, device="cuda") - 0.5).half()       # shadow of the trailing matrix
S32 = torch.zeros(KB, N, device="cuda")
S16 = (torch.rand(KB, N, device="cuda") - 0.5).half()
C = torch.rand(D, N, device="cuda")
H = torch.zeros(D, N, device="cuda", dtype=torch.float16)


def timeit(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for mode in ("2cta", "1cta"):
    if mode == "1cta":
        os.environ["MPQR_GEMM_1CTA"] = "1"
    t_tn = timeit(lambda: pkg.check(L.mpqr_gemm_tn_device(W.data_ptr(), KB, A16.data_ptr(), N, S32.data_ptr(), N, KB, N, D, 0, st)))
    t_nn = timeit(lambda: pkg.check(L.mpqr_gemm_nn_device(W.data_ptr(), KB, S16.data_ptr(), N, C.data_ptr(), N, H.data_ptr(), N, D, N, KB, 0, st)))
    fl = 2.0 * D * N * KB
    print(f"{mode}: TN S[{KB}x{N}] = W^T A (K={D}): {t_tn:.3f} ms {fl / t_tn / 1e9:.0f} TFLOP/s | "
          f"NN C[{D}x{N}] -= Y S (K={KB}): {t_nn:.3f} ms {fl / t_nn / 1e9:.0f} TFLOP/s, {10.0 * D * N / t_nn / 1e6:.0f} GB/s algorithmic", flush=True)
