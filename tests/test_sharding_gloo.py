"""N > 1 host logic on CPU: world_size-2 gloo processes shard a corpus, search their shard
(with the oracle standing in for the GPU kernel), all-gather [nq][k] lists and merge -- the result
must equal the unsharded search.  Mirrors bench.py's multi-GPU data flow."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from myscaledb_b200.sharding import assign_parts, shard_range
    for total, world, align in [(10_000_000, 8, 250_000), (1001, 3, 1), (7, 8, 1), (10_000_000, 1, 250_000)]:
        rs = [shard_range(total, world, r, align) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    a = assign_parts([10, 1, 7, 3, 3, 8], 2)
    assert sorted(sum(a, [])) == list(range(6))
    assert abs(sum([10, 1, 7, 3, 3, 8][i] for i in a[0]) - 16) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle as orc
    from myscaledb_b200.sharding import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((5000, 32)).astype(np.float32)
    x = rng.standard_normal((9, 32)).astype(np.float32)
    k = 7
    lo, hi = shard_range(len(y), world, rank)
    d, i = orc.knn_flat(orc.IP, x, y[lo:hi], k)
    i = np.where(i >= 0, i + lo, -1)
    gd = [torch.empty((9, k)) for _ in range(world)]
    gi = [torch.empty((9, k), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gd, torch.from_numpy(d))
    dist.all_gather(gi, torch.from_numpy(i))
    if rank == 0:
        D = torch.stack(gd).numpy(); I = torch.stack(gi).numpy()
        out_i = np.empty((9, k), np.int64)
        for qi in range(9):
            sc = D[:, qi].reshape(-1); lab = I[:, qi].reshape(-1)
            order = np.lexsort((lab, -sc))[:k]
            out_i[qi] = lab[order]
        _, ref = orc.knn_flat(orc.IP, x, y, k)
        q.put(bool((out_i == ref).all()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    ok = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert ok and all(p.exitcode == 0 for p in ps)
