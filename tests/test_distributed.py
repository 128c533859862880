"""N > 1 path on CPU: world_size-2 (and 3) gloo process groups exercising the sharding and the single
collective of the hot path (final cost / argmin gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygent_b200 import distributed as pgd


def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 7, 16384, 65536, 65537):
        for w in (1, 2, 3, 8):
            b = [pgd.shard_bounds(total, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, costs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.as_tensor(costs)
        mine, lo = pgd.shard(full)
        best = pgd.gather_best(mine, offset=lo)
        allc = pgd.gather_costs(mine)
        empty = pgd.gather_best(mine[:0] if rank == 0 else mine, offset=lo)  # a rank with nothing to contribute
        q.put((rank, best, allc.numpy().tolist(), empty))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_argmin_gather_matches_host_argmin(world):
    rng = np.random.default_rng(world)
    costs = rng.uniform(1, 100, 1001)
    costs[[17, 640]] = 0.5  # a tie: the lower global index must win, like np.argmin
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, costs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, allc, empty in res:
        assert best[1] == int(np.argmin(costs)) == 17 and best[0] == 0.5
        assert best[2] == 0  # owner rank of index 17
        assert allc == costs.tolist()
        lo, hi = pgd.shard_bounds(len(costs), 0, world)
        rest = costs.copy()
        rest[lo:hi] = np.inf
        assert empty[1] == int(np.argmin(rest))
