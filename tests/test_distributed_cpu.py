"""world_size-2 gloo test of the N>1 host logic: member sharding + the final gather (the only
collective of the path).  Compute is replaced by a deterministic function of the member index so the
test needs no GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decipher_b200.distributed import PerfGather, gather_to_rank0, lpt_assign, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None, None] * torch.ones(1, 3, 2, dtype=torch.float64)
    full = gather_to_rank0(local, n_total)
    # the preallocated equal-block gather used once per run by bench.py / ensemble drivers: two calls reuse the buffers
    pg = PerfGather(4, (3, 2), world, rank, torch.device("cpu"))
    for rep in range(2):
        blk = (torch.arange(4, dtype=torch.float64) + 100 * rank + 1000 * rep)[:, None, None] * torch.ones(1, 3, 2, dtype=torch.float64)
        got = pg(blk)
        if rank == 0:
            assert got.shape == (world, 4, 3, 2) and got.data_ptr() == pg.out.data_ptr()
            for r in range(world):
                assert torch.equal(got[r, :, 0, 0], torch.arange(4, dtype=torch.float64) + 100 * r + 1000 * rep)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    for n, w in [(10, 1), (10, 3), (1000000, 8), (7, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_total = 11                      # ragged: 6 + 5
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    full = q.get(timeout=120)
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (n_total, 3, 2)
    assert np.array_equal(full[:, 0, 0], np.arange(n_total, dtype=np.float64))


def test_lpt_assignment_balances_catchments():
    rng = np.random.default_rng(0)
    costs = rng.integers(100, 1000, 200).astype(float)
    owner = lpt_assign(costs, 8)
    load = np.bincount(owner, weights=costs, minlength=8)
    assert load.max() / load.mean() < 1.02
