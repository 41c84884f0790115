"""world_size-2 CPU test (gloo) of the row-sharded H.v scheme: contiguous row blocks chosen by the
library's partition rule (sqmc_b200_partition_rows, pure host code), FULL rows per shard, all-gather
of the vector slices, no reduction of the result -- the data flow the CUDA path runs over NCCL.
The arithmetic here is numpy on the oracle's matrix; the CUDA kernels are covered by -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import ctypes as C
    from oracle import oracle as O
    from sqmc_b200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = O.System.heg(3, 0.5, 14, 7, 1.49)
    r = S.hci(1e-3, n_states=1, max_iters=1)
    cnt, idx, val = S.build_upper(r["up"], r["dn"])
    A = O.upper_to_scipy(cnt, idx, val).tocsr()
    n = A.shape[0]
    prefix = np.zeros(n + 1, dtype=np.int64)
    prefix[1:] = np.cumsum(np.diff(A.indptr))
    starts = np.zeros(world + 1, dtype=np.int64)
    L = _lib.load()
    assert L.sqmc_b200_partition_rows(prefix.ctypes.data_as(C.c_void_p), n, world, starts.ctypes.data_as(C.c_void_p)) == 0
    r0, r1 = int(starts[rank]), int(starts[rank + 1])
    Aloc = A[r0:r1]                                  # full rows of this shard, global column ids
    x_full = np.random.default_rng(11).uniform(-1, 1, n)
    x_loc = torch.from_numpy(x_full[r0:r1].copy())   # each rank owns its slice of the Krylov vector
    pieces = [torch.zeros(int(starts[k + 1] - starts[k]), dtype=torch.float64) for k in range(world)]
    dist.all_gather(pieces, x_loc) if len(set(p.numel() for p in pieces)) == 1 else _allgatherv(pieces, x_loc, rank, world)
    xg = torch.cat(pieces).numpy()
    y_loc = Aloc @ xg
    # partial dot for the Krylov column, completed by a small all-reduce
    part = torch.tensor([float(np.dot(x_full[r0:r1], y_loc))], dtype=torch.float64)
    dist.all_reduce(part)
    y_ref = O.matvec_upper(cnt, idx, val, x_full)
    ok = np.allclose(xg, x_full) and np.allclose(y_loc, y_ref[r0:r1], rtol=0, atol=1e-12) and abs(part.item() - np.dot(x_full, y_ref)) < 1e-10
    q.put((rank, bool(ok), r0, r1))
    dist.destroy_process_group()


def _allgatherv(pieces, x_loc, rank, world):
    # in-place all-gather of unequal blocks = one broadcast per owner (what allgather_rows does with ncclBroadcast)
    for k in range(world):
        if k == rank:
            pieces[k].copy_(x_loc)
        dist.broadcast(pieces[k], src=k)


def test_row_sharded_matvec_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    spans = sorted((r0, r1) for _, _, r0, r1 in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 277
