"""Worker for the multi-GPU parity test (launched with torchrun, one process per GPU): factors a seeded matrix with
the column-block-cyclic look-ahead driver and checks the gathered factor against the oracle (LAPACK FP64 for the
larger shapes) and against the single-GPU driver."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixedprecisionblockqr_b200 as pkg  # noqa: E402


def sampled_backward_error(A0, P, r, k=16):
    """||(A - QR) X||_F / (||A||_F sqrt(k)), Gaussian X, FP64, from the packed factor (GPU tensors)."""
    m, n = A0.shape
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, k, device="cuda", dtype=torch.float64, generator=g)
    AX = A0.double() @ X
    kmax = min(m, n)
    Z = torch.triu(P[:m].double()) @ X
    for lam in range(((kmax - 1) // r) * r, -1, -r):
        pw = min(r, kmax - lam)
        Y = torch.tril(P[lam + 1:m + 1, lam:lam + pw].double())
        Tinv = torch.triu(Y.T @ Y, 1) + 0.5 * torch.eye(pw, device="cuda", dtype=torch.float64)
        Z[lam:] -= Y @ torch.linalg.solve_triangular(Tinv, Y.T @ Z[lam:], upper=True)
    return (torch.linalg.norm(AX - Z) / (torch.linalg.norm(A0.double()) * k ** 0.5)).item()


def main():
    m, n, r, nb = [int(x) for x in sys.argv[1:5]]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    uid = [pkg.mg_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    plan = pkg.MultiGpuBlockQR(m, n, r, nb, rank, world, uid[0])
    gcols = pkg.mg_layout_global_cols(n, plan.nb, rank, world)
    nloc = plan.local_cols
    lda = (max(nloc, 8) + 7) // 8 * 8
    st = torch.cuda.current_stream().cuda_stream
    # every rank generates the full seeded matrix on its GPU and keeps its own columns
    full = torch.zeros(m + 1, n, device="cuda")
    pkg.fill_uniform(full.data_ptr(), n, n, 0, m, 0, n, 4242, st)
    A = torch.zeros(m + 1, lda, device="cuda")
    A[:, :nloc] = full[:, torch.from_numpy(gcols).cuda()]
    plan.factor(A.data_ptr(), lda, st)
    torch.cuda.synchronize()
    # gather the packed factor on rank 0
    parts = [None] * world
    dist.all_gather_object(parts, (gcols, A[:, :nloc].cpu().numpy()))
    ok = True
    if rank == 0:
        P = np.zeros((m + 1, n), np.float32)
        for gc, blk in parts:
            P[:, gc] = blk
        # Reference: the ORACLE's packed factor where it finishes in seconds (the restatement of h_block_qr,
        # Cuda/qr.cu:1275), else |R| from FP64 LAPACK; plus the single-GPU driver on the same input.
        A_host = full[:m].cpu().numpy()
        if m * n * n <= 3e9:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import oracle
            Pref, _ = oracle.block_qr(A_host, r, want_q=False)
            Rref = np.triu(Pref[:m])
            ref_name = "oracle"
        else:
            Rref = np.linalg.qr(A_host.astype(np.float64), mode="r")
            ref_name = "LAPACK fp64"
        Rm = np.triu(P[:m])[:Rref.shape[0]]
        d = np.abs(np.abs(Rm) - np.abs(Rref)).max() / np.abs(Rref).max()
        single = pkg.BlockQR(m, n, r, nb=plan.nb, precision="fp16")
        ref = full.clone()
        single.factor(ref.data_ptr(), n, st)
        torch.cuda.synchronize()
        Rs = np.triu(np.abs(ref.cpu().numpy()[:m]))[:Rref.shape[0]]
        ds = np.abs(np.abs(Rm) - Rs).max() / Rs.max()
        be = sampled_backward_error(full[:m], torch.from_numpy(P).cuda(), plan.r)
        bes = sampled_backward_error(full[:m], ref, plan.r)
        print(f"mg({world}) {m}x{n} r={r} nb={plan.nb}: |R| rel max diff vs {ref_name} {d:.3e}, vs single GPU {ds:.3e}; "
              f"sampled backward error mg {be:.3e} single {bes:.3e}")
        eps = 2.0 ** -11
        ok = d <= 10 * eps and ds <= 10 * eps and be <= 5.5 * eps
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
