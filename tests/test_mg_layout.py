"""CPU: host-side logic of the multi-GPU path — column-block-cyclic layout (pure C helpers of
libmpqr.so) and the rendezvous / unique-id exchange / max-over-ranks plumbing bench.py uses,
run with world_size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mixedprecisionblockqr_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,nb,P", [(32768, 1024, 8), (1000, 128, 4), (2048, 256, 2), (100, 64, 3), (5, 8, 2), (4096, 1024, 1)])
def test_layout_partitions_columns(n, nb, P):
    seen = []
    for rank in range(P):
        g = pkg.mg_layout_global_cols(n, nb, rank, P)
        assert len(g) == pkg.mg_layout_local_cols(n, nb, rank, P)
        assert np.all(np.diff(g) > 0)                       # local order = global order
        assert np.all((g // nb) % P == rank)                # block b lives on rank b % P
        seen.append(g)
    allc = np.sort(np.concatenate(seen))
    assert np.array_equal(allc, np.arange(n))               # every column owned exactly once


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # unique-id style exchange: rank 0 creates 128 opaque bytes, everybody must receive the same
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g = pkg.mg_layout_global_cols(n, nb, rank, world)
    # each rank "generates" its shard from the global (row, col) hash and the shards reassemble the matrix
    import oracle
    full = oracle.uniform_matrix(16, n, 5)
    parts = [None] * world
    dist.all_gather_object(parts, (g, full[:, g]))
    rebuilt = np.zeros_like(full)
    for gc, blk in parts:
        rebuilt[:, gc] = blk
    # max-over-ranks timing reduction as in bench.py
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, uid[0] == bytes(range(128)), bool(np.array_equal(rebuilt, full)), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, 128, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid_ok, rebuilt_ok, tmax in res:
        assert uid_ok and rebuilt_ok and tmax == 11.0


def _tsqr_worker(rank, world, port, m, n, q):
    """The exchange structure of mpqr_mg_tsqr_device over gloo, with the oracle's ts_qr restatement as the leaf
    factoriser: row-block leaves -> all_gather of the n x n R factors -> the same stack factored on every rank
    -> local thin-Q product.  Checks that the structure reproduces A = Q R with the same R on all ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    A = oracle.uniform_matrix(m, n, 77).astype(np.float64)
    rows = np.array_split(np.arange(m), world)[rank]        # 1-D row-block layout
    Qp, Rp = oracle.tsqr(A[rows], 4)                        # leaf (python/ca_qr.py:25-43 restated)
    Qp = Qp[:, :n]
    gathered = [torch.zeros(n, n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(np.ascontiguousarray(Rp[:n])))
    stack = torch.cat(gathered).numpy()
    Qs, R = np.linalg.qr(stack)                             # redundant on every rank, same input => same R
    Qloc = Qp[: len(rows)] @ Qs[rank * n:(rank + 1) * n]
    Rs = [torch.zeros(n, n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(Rs, torch.from_numpy(np.ascontiguousarray(R)))
    same_r = all(torch.equal(Rs[0], x) for x in Rs)
    Qall = [None] * world
    dist.all_gather_object(Qall, (rows, Qloc))
    Q = np.zeros((m, n))
    for rr, blk in Qall:
        Q[rr] = blk
    err = np.linalg.norm(A - Q @ R) / np.linalg.norm(A)
    orth = np.linalg.norm(Q.T @ Q - np.eye(n))
    q.put((rank, same_r, float(err), float(orth)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_tsqr_structure():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tsqr_worker, args=(r, 2, port, 640, 24, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_r, err, orth in res:
        assert same_r and err < 1e-13 and orth < 1e-12
