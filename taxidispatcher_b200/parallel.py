"""Multi-GPU partitioning of the parts of the hot path that shard naturally.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  No data-path
collective inside any kernel -- the only exchange is the reference's own: gather the per-shard
survivors, then re-run the greedy scan on the concatenation (findpool.c:149-172).

  pool      the 8 LOGICAL shards of pool_n.c:226-229 are kept whatever the GPU count (the result
            depends on them: each shard dedups before the merge); rank r runs a contiguous block
            of logical shards; one all_gather of fixed-capacity survivor buffers; every rank (or
            rank 0 only) merges in shard order.
  cost      contiguous cab-row blocks per rank (rows_for_rank); no exchange.
  assign    independent instances (split.py ranges, per-minute batches) round-robin over ranks.
  LCM       does not shard (sequential chain): replicas only.

`compute_shard` / `merge` are injected so that the host logic can be exercised on CPU with the
oracle standing in for the device (tests only); the product default is the CUDA engine.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

REC_W = 9
REF_SHARDS = 8  # pool_n.c:13 MAX_THREAD


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shards_for_rank(rank: int, world_size: int, n_shards: int = REF_SHARDS) -> List[int]:
    """Logical shards owned by `rank`: a CONTIGUOUS block (balanced to within one shard), so that one
    td_pool_find_shards call serves all of a rank's shards."""
    lo, hi = rows_for_rank(n_shards, rank, world_size)
    return list(range(lo, hi))


def rows_for_rank(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous cab-row block [lo, hi) of the cost matrix for `rank` (balanced to within one row)."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def instances_for_rank(n_instances: int, rank: int, world_size: int) -> List[int]:
    return list(range(rank, n_instances, world_size))


def _device_for_collectives() -> torch.device:
    backend = dist.get_backend() if dist.is_initialized() else "gloo"
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def gather_shard_plans(local: Sequence[Tuple[int, np.ndarray]], n: int, n_shards: int = REF_SHARDS) -> List[np.ndarray]:
    """local: [(shard index, plans [m,9])] computed by this rank.  Returns the plans of ALL logical
    shards in shard order on every rank.  One all_gather of a fixed-capacity int32 buffer:
    per shard a count followed by cap x 9 ints, cap = n // 2 + 1 (survivors are customer-disjoint)."""
    rank, w = world()
    cap = n // 2 + 1
    slots = (n_shards + w - 1) // w          # shards per rank, padded
    stride = 2 + cap * REC_W                  # [shard id, count, rows...]
    buf = np.full((slots, stride), -1, dtype=np.int32)
    for slot, (sh, plans) in enumerate(local):
        plans = np.asarray(plans, dtype=np.int32).reshape(-1, REC_W)
        if len(plans) > cap:
            raise ValueError("shard %d returned %d plans, capacity %d" % (sh, len(plans), cap))
        buf[slot, 0] = sh
        buf[slot, 1] = len(plans)
        buf[slot, 2: 2 + plans.size] = plans.reshape(-1)
    if w == 1:
        gathered = [buf]
    else:
        dev = _device_for_collectives()
        mine = torch.from_numpy(buf).to(dev)
        outs = [torch.empty_like(mine) for _ in range(w)]
        dist.all_gather(outs, mine)
        gathered = [o.cpu().numpy() for o in outs]
    by_shard = {}
    for g in gathered:
        for row in g:
            if row[0] >= 0:
                by_shard[int(row[0])] = row[2: 2 + int(row[1]) * REC_W].reshape(-1, REC_W).copy()
    return [by_shard[s] for s in range(n_shards)]


def cost_matrix_sharded(distances, cab_to, cust_from, fill: int = 250000, cutoff: Optional[int] = None,
                        gather: bool = False, compute_rows: Optional[Callable] = None):
    """calculate_cost (split.py:123-136) with the cab rows split across the ranks of the current process group: rank r
    builds rows rows_for_rank(n, r, world) of the padded n x n matrix on its own device -- no exchange, the loop of
    split.py:129-134 carries no state from row to row.  Returns (n, (lo, hi), block) with block a (hi - lo) x n int32
    device tensor; gather=True all-gathers the blocks (one all_gather_into_tensor per call, padded to equal heights) and
    returns the whole matrix on every rank instead.  compute_rows(lo, hi) -> array stands in for the device in CPU
    tests."""
    rank, w = world()
    cab = np.ascontiguousarray(np.asarray(cab_to, dtype=np.int32))
    cust = np.ascontiguousarray(np.asarray(cust_from, dtype=np.int32))
    n = max(len(cab), len(cust))
    lo, hi = rows_for_rank(n, rank, w)
    if compute_rows is not None:
        block = torch.as_tensor(np.asarray(compute_rows(lo, hi), dtype=np.int32).reshape(hi - lo, n))
    else:
        from . import dispatch
        eng = dispatch.engine()
        block = eng.cost_matrix(dispatch._h2d_i32(distances), dispatch._h2d_i32(cab), dispatch._h2d_i32(cust), fill, cutoff,
                                rows=(lo, hi))
    if not gather or w == 1:
        return n, (lo, hi), block
    height = (n + w - 1) // w                       # blocks differ by at most one row: pad to the common height
    dev = _device_for_collectives()
    mine = torch.zeros((height, n), dtype=torch.int32, device=dev)
    mine[: hi - lo] = block.to(dev)
    full = torch.empty((w * height, n), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(full, mine)
    parts = []
    for r in range(w):
        rlo, rhi = rows_for_rank(n, r, w)
        parts.append(full[r * height: r * height + (rhi - rlo)])
    return n, (0, n), torch.cat(parts, dim=0)


_SLOT_SHARD = {}


def _find_pool_sharded_device(dem: np.ndarray, dist_table, pool_size: int, n_shards: int, rank: int, w: int, mine: List[int]):
    """Device-resident product path (NCCL): this rank's block of shards in one asynchronous device call into HEADED blocks
    (survivors + their count + the shard's counters, td_pool_find_shards_headed), ONE all_gather into every rank's merge
    input, merge on the device, ONE read-back.  No stats all_reduce, no second collective, no host round trip before the
    merge.  Returns None (on every rank together: all ranks see the same headers) when some rank's record list
    overflowed -- the caller then takes the synchronous path, which sizes the list for the next call."""
    from . import dispatch
    eng = dispatch.engine()
    dev = eng.device
    n = dem.shape[0]
    cap = n // 2 + 1
    slots = (n_shards + w - 1) // w
    dist_np = np.ascontiguousarray(np.asarray(dist_table, dtype=np.int32))

    def result(plans, counts, ev, fe):
        kept = [0] * n_shards
        for r in range(w):
            for k_, sh in enumerate(shards_for_rank(r, w, n_shards)):
                kept[sh] = int(counts[r * slots + k_])
        return plans, {"evaluated": int(ev.sum()), "feasible": int(fe.sum()), "kept_per_shard": kept, "kept": len(plans)}

    # steady state of a repeated shape: H2D, this rank's shards, the all_gather, the merge and the read-back are ONE graph
    # launch (dispatch.PoolJobGraph).  A rank whose graph was invalidated falls through to the piecewise path below: both
    # issue the same single all_gather, so the ranks stay matched.
    gkey = ("sharded", w, n_shards, n, dist_np.shape[0], pool_size, dev.index)
    job = dispatch._POOL_GRAPHS.get(gkey)
    if isinstance(job, dispatch.PoolJobGraph) and not job.valid():
        dispatch._POOL_GRAPHS.pop(gkey, None)
        job = None
    if isinstance(job, dispatch.PoolJobGraph):
        plans, counts, ev, fe = job.run(dem, dist_np)
        if int(counts.min()) < 0:
            dispatch._POOL_GRAPHS.pop(gkey, None)
            return None
        return result(plans, counts, ev, fe)
    dem_d, dist_d = dispatch._h2d_i32_many([dem, dist_np])
    key = (w, n_shards, n, dev.index)
    if key not in _SLOT_SHARD:      # buffers + the logical shard of every (rank, slot); padding slots keep count 0 for ever
        ids = []
        for r in range(w):
            sh = shards_for_rank(r, w, n_shards)
            ids += sh + [0] * (slots - len(sh))
        _SLOT_SHARD[key] = (torch.tensor(ids, dtype=torch.int32, device=dev), ids,
                            torch.zeros((slots, cap + 1, REC_W), dtype=torch.int32, device=dev),
                            torch.empty((w * slots, cap + 1, REC_W), dtype=torch.int32, device=dev))
    slot_shard, ids, blocks, all_blocks = _SLOT_SHARD[key]
    if mine:
        eng.pool_find_shards_headed(dem_d, dist_d, pool_size, mine[0], len(mine), n_shards, out=blocks[: len(mine)])
    dist.all_gather_into_tensor(all_blocks, blocks)
    plans, counts, ev, fe = eng.pool_merge_headed_packed(all_blocks, slot_shard, n, pool_size)
    if int(counts.min()) < 0:
        dispatch._POOL_GRAPHS.pop(gkey, None)
        return None
    if dispatch._pool_graphs_enabled() and n > 0 and dist_np.ndim == 2:
        seen = dispatch._POOL_GRAPHS.get(gkey, 0) + 1          # second success of the shape: capture the job (every rank does)
        dispatch._POOL_GRAPHS[gkey] = seen if seen < 2 else dispatch._capture_pool_job(
            eng, n, dist_np.shape[0], pool_size, n_shards, mine[0] if mine else 0, len(mine),
            gather=(dist.all_gather_into_tensor, w), slots=slots, slot_shard=slot_shard)
    return result(plans, counts, ev, fe)


def find_pool_sharded(demand, dist_table, pool_size: int, n_shards: int = REF_SHARDS,
                      compute_shard: Optional[Callable] = None, merge: Optional[Callable] = None):
    """`findpool` over the ranks of the current process group: returns (merged plans, stats) on every
    rank.  stats carry the sums over all logical shards (evaluated, feasible) and kept_per_shard."""
    rank, w = world()
    dem = np.asarray(demand, dtype=np.int32).reshape(-1, 5)
    n = dem.shape[0]
    mine = shards_for_rank(rank, w, n_shards)
    if compute_shard is None and merge is None and w > 1 and dist.get_backend() == "nccl" and n_shards <= 64 * w:
        fast = _find_pool_sharded_device(dem, dist_table, pool_size, n_shards, rank, w, mine)
        if fast is not None:
            return fast
    if compute_shard is None:
        from . import dispatch                      # product path: one device call for the whole block
        results = dispatch.find_pool_block(dem, dist_table, pool_size, mine[0], len(mine), n_shards) if mine else []
    else:
        results = [compute_shard(dem, dist_table, pool_size, sh, n_shards) for sh in mine]
    if merge is None:
        from . import dispatch
        merge = dispatch.pool_merge
    local = []
    ev = fe = 0
    for sh, (plans, st) in zip(mine, results):
        local.append((sh, plans))
        ev += int(st["evaluated"])
        fe += int(st["feasible"])
    all_plans = gather_shard_plans(local, n, n_shards)
    if w > 1:
        t = torch.tensor([ev, fe], dtype=torch.int64, device=_device_for_collectives())
        dist.all_reduce(t)
        ev, fe = int(t[0]), int(t[1])
    merged = merge(all_plans, n, pool_size)
    return merged, {"evaluated": ev, "feasible": fe, "kept_per_shard": [len(p) for p in all_plans], "kept": len(merged)}
