"""Quick device-resident timing of the factorisation (CUDA events) + sampled backward error."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mixedprecisionblockqr_b200 as pkg


def sampled_backward_error(A0, P, r=128, k=16):
    """|| (A - Q R) X ||_F / (||A||_F sqrt(k)) with Gaussian X, evaluated in FP64 from the packed factor."""
    m, n = A0.shape
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, k, device="cuda", dtype=torch.float64, generator=g)
    AX = A0.double() @ X
    kmax = min(m, n)
    R = torch.triu(P[:m].double())
    Z = R @ X                      # m x k
    for lam in range(((kmax - 1) // r) * r, -1, -r):
        pw = min(r, kmax - lam)
        Y = torch.tril(P[lam + 1:m + 1, lam:lam + pw].double())   # shifted storage -> D x pw unit vectors
        G = Y.T @ Y
        Tinv = torch.triu(G, 1) + 0.5 * torch.eye(pw, device="cuda", dtype=torch.float64)
        # Q_p = I - Y T Y^T ; Z[lam:] <- Q_p Z[lam:]
        Z[lam:] -= Y @ torch.linalg.solve_triangular(Tinv, Y.T @ Z[lam:], upper=True)
    return (torch.linalg.norm(AX - Z) / (torch.linalg.norm(A0.double()) * k ** 0.5)).item()


def run(m, n, r, prec, nb=0, reps=3, check=True):
    lda = (n + 7) // 8 * 8
    A = torch.zeros(m + 1, lda, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    plan = pkg.BlockQR(m, n, r, nb=nb, precision=prec)
    times = []
    for it in range(reps):
        pkg.fill_uniform(A.data_ptr(), lda, n, 0, m, 0, n, 1234, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.factor(A.data_ptr(), lda, st)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    F = pkg.householder_flops(m, n)
    best = min(times)
    be = float("nan")
    if check:
        A0 = torch.zeros(m, lda, device="cuda")
        pkg.fill_uniform(A0.data_ptr(), lda, n, 0, m, 0, n, 1234, st)
        be = sampled_backward_error(A0[:, :n], A[:, :n], r=plan.r)
    print(f"{m}x{n} r={plan.r} nb={plan.nb} {prec}: {best:.2f} ms  {F / best / 1e9:.2f} TFLOP/s  launches={plan.last_launches} "
          f"bwd_err~{be:.2e}  all={['%.1f' % t for t in times]}", flush=True)
    if os.environ.get("MPQR_TRACE") == "1":
        pkg.lib().mpqr_debug_dump_trace(plan._h)
    if os.environ.get("PROFILE") == "1":
        import ctypes
        L = pkg.lib()
        L.mpqr_set_profiling(plan._h, 1)
        pkg.fill_uniform(A.data_ptr(), lda, n, 0, m, 0, n, 1234, st)
        plan.factor(A.data_ptr(), lda, st)
        torch.cuda.synchronize()
        names = ["panel(all)", "gemm_tn", "gemm_nn", "cast", "panel:block", "panel:S", "panel:G/T/W", "panel:U"]
        out = []
        for c in range(8):
            ms, cnt, fl, by = ctypes.c_double(), ctypes.c_long(), ctypes.c_double(), ctypes.c_double()
            L.mpqr_get_profile(plan._h, c, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(fl), ctypes.byref(by))
            out.append(f"{names[c]}={ms.value:.2f}ms/{cnt.value}")
        print("   profile: " + "  ".join(out), flush=True)
        L.mpqr_set_profiling(plan._h, 0)
    plan.close()
    del A
    torch.cuda.empty_cache()


if __name__ == "__main__":
    reps = int(os.environ.get("REPS", "3")); chk = os.environ.get("CHECK", "1") == "1"
    cfgs = sys.argv[1:] or ["2048,2048,32,fp16", "2048,2048,32,fp32", "4096,16384,64,fp16", "8192,8192,128,fp16", "16384,16384,128,fp16"]
    for c in cfgs:
        parts = c.split(",")
        m, n, r, prec = int(parts[0]), int(parts[1]), int(parts[2]), parts[3]
        nb = int(parts[4]) if len(parts) > 4 else 0
        run(m, n, r, prec, nb=nb, reps=reps, check=chk)
