"""CPU suite: the multi-GPU host logic (shard ownership, gather layout, merge order) under a real
world_size-2 gloo process group.  The oracle stands in for the device here -- this tests plumbing,
not kernels (those are covered by -m gpu)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_helpers():
    from taxidispatcher_b200 import parallel as P
    for w in (1, 2, 3, 4, 8):
        owned = sorted(s for r in range(w) for s in P.shards_for_rank(r, w))
        assert owned == list(range(8))
        rows = [P.rows_for_rank(20000, r, w) for r in range(w)]
        assert rows[0][0] == 0 and rows[-1][1] == 20000
        assert all(rows[i][1] == rows[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in rows) - min(b - a for a, b in rows) <= 1
        inst = sorted(i for r in range(w) for i in P.instances_for_rank(5, r, w))
        assert inst == list(range(5))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gen_inputs as g, pool_ref
    from taxidispatcher_b200 import parallel as P
    dem = g.pool_demand(150, seed=21)
    d = g.stand_distances(51)
    for k in (2, 4):
        merged, st = P.find_pool_sharded(dem, d, k, compute_shard=pool_ref.find, merge=pool_ref.merge)
        q.put((rank, k, merged.tolist(), st))
    dist.barrier()
    dist.destroy_process_group()


def _cost_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cost_ref, gen_inputs as g
    from taxidispatcher_b200 import parallel as P
    d, cab_to, cust_from = g.config1b()          # 190 cabs x 200 customers: 10 dummy rows

    def rows(lo, hi):                              # the literal loops of split.py:123-136 stand in for the device
        return cost_ref.calculate_cost_np(d, cab_to, cust_from)[1][lo:hi]

    n, (lo, hi), block = P.cost_matrix_sharded(d, cab_to, cust_from, compute_rows=rows)
    n2, span, full = P.cost_matrix_sharded(d, cab_to, cust_from, gather=True, compute_rows=rows)
    q.put((rank, n, lo, hi, block.numpy().tolist(), span, full.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_cost_matrix_sharded_gloo(world):
    """cab rows split over the ranks (north_star), blocks of unequal height, optional all-gather of the whole matrix"""
    from oracle import cost_ref, gen_inputs as g
    from taxidispatcher_b200 import parallel as P
    d, cab_to, cust_from = g.config1b()
    n_ref, ref = cost_ref.calculate_cost_np(d, cab_to, cust_from)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cost_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, lo, hi, block, span, full in got:
        assert n == n_ref and (lo, hi) == P.rows_for_rank(n, rank, world)
        assert block == ref[lo:hi].tolist()
        assert tuple(span) == (0, n) and full == ref.tolist()


@pytest.mark.parametrize("world", [2, 3])
def test_find_pool_sharded_gloo(world):
    from oracle import gen_inputs as g, pool_ref
    dem = g.pool_demand(150, seed=21)
    d = g.stand_distances(51)
    expect = {}
    for k in (2, 4):
        shards = [pool_ref.find(dem, d, k, sh) for sh in range(8)]
        expect[k] = (pool_ref.merge([p for p, _ in shards], 150, k).tolist(),
                     sum(s["evaluated"] for _, s in shards), [len(p) for p, _ in shards])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world * 2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, k, merged, st in got:
        assert merged == expect[k][0], (rank, k)
        assert st["evaluated"] == expect[k][1] and st["kept_per_shard"] == expect[k][2]
