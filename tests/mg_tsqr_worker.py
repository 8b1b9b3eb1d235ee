"""Worker for the multi-GPU TSQR parity test (torchrun, one process per GPU): row-block TSQR through
mpqr_mg_tsqr_device against LAPACK on the gathered matrix."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixedprecisionblockqr_b200 as pkg  # noqa: E402


def main():
    m, n = [int(x) for x in sys.argv[1:3]]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    uid = [pkg.mg_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    plan = pkg.MultiGpuTSQR(rank, world, uid[0])
    st = torch.cuda.current_stream().cuda_stream
    bounds = np.linspace(0, m, world + 1).astype(int)   # uneven row blocks are fine: only the R factors travel
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    mloc = r1 - r0
    A = torch.zeros(mloc, n, device="cuda")
    pkg.fill_uniform(A.data_ptr(), n, n, r0, mloc, 0, n, 99, st)   # this rank's rows of the seeded matrix
    Q = torch.zeros(mloc, n, device="cuda")
    R = torch.zeros(n, n, device="cuda")
    plan.factor(A.data_ptr(), n, mloc, n, Q.data_ptr(), n, R.data_ptr(), n, st)
    R2 = torch.zeros(n, n, device="cuda")
    plan.factor(A.data_ptr(), n, mloc, n, None, n, R2.data_ptr(), n, st)   # R only
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (A.cpu().numpy(), Q.cpu().numpy(), R.cpu().numpy(), R2.cpu().numpy()))
    ok = True
    if rank == 0:
        Af = np.concatenate([p[0] for p in parts]).astype(np.float64)
        Qf = np.concatenate([p[1] for p in parts]).astype(np.float64)
        Rr = parts[0][2].astype(np.float64)
        # rank 0's R is broadcast: identical everywhere; a second call agrees to rounding (atomics in the reductions)
        same = all(np.array_equal(parts[0][2], p[2]) for p in parts) and all(
            np.abs(p[3] - parts[0][2]).max() <= 2e-6 * np.abs(parts[0][2]).max() for p in parts)
        _, Rl = np.linalg.qr(Af)
        d = np.abs(np.abs(Rr) - np.abs(Rl)).max() / np.abs(Rl).max()
        be = np.linalg.norm(Af - Qf @ Rr) / np.linalg.norm(Af)
        orth = np.linalg.norm(Qf.T @ Qf - np.eye(n))
        print(f"mg tsqr({world}) {m}x{n}: R identical on ranks {same}; |R| vs LAPACK {d:.2e}; backward {be:.2e}; orth {orth:.2e}")
        ok = same and d <= 2e-5 and be <= 5e-6 and orth <= 5e-5 and np.allclose(np.triu(Rr), Rr)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    plan.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
