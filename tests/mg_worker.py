"""Worker for the 2-GPU parity test (launched with torchrun, one process per GPU): factors a
seeded matrix with the column-block-cyclic driver and checks it against the single-GPU driver."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mixedprecisionblockqr_b200 as pkg  # noqa: E402


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
        single = pkg.BlockQR(m, n, r, nb=plan.nb, precision="fp16")
        ref = full.clone()
        single.factor(ref.data_ptr(), n, st)
        torch.cuda.synchronize()
        Pref = ref.cpu().numpy()
        d = np.abs(P - Pref).max() / np.abs(Pref).max()
        print(f"mg({world}) vs single-GPU packed factor: rel max diff {d:.3e}")
        ok = d <= 1e-5   # same kernels, same operand order -> identical up to split-K atomics order
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
