"""calculate_cost restated.  TEST INFRASTRUCTURE ONLY.

Reference: split.py:123-136 (= greedy_opt.py:86-99), cutoff variant simulate.py:17-33 /
Simulator.java:493-520, id-indexed variant with fill n*n procedure.py:6-12.
Parity status: PINNED by KAT A3 (procedure.py:32-51 -> [[3,3,0,2],[1,1,2,4],[5,5,2,0],[16]*4]).
"""
from __future__ import annotations

import numpy as np

BIG_COST = 250000


def calculate_cost(distances, demand, cabs, fill=BIG_COST, cutoff=None):
    """Literal double loop of split.py:123-136; `cutoff` adds simulate.py:27's `< DROP_TIME` test.

    demand / cabs: lists of (id, from, to).  Positional indexing (c_idx, d_idx), like split.py.
    Returns (n, cost) with cost a list of n rows.  n == 0 -> (0, 0) as simulate.py:21.
    """
    n = len(cabs) if len(cabs) > len(demand) else len(demand)
    if n == 0:
        return 0, 0
    cost = [[fill for _ in range(n)] for _ in range(n)]
    for c_idx, (_c_id, _c_frm, c_to) in enumerate(cabs):
        for d_idx, (_d_id, d_frm, _d_to) in enumerate(demand):
            d = int(distances[c_to][d_frm])
            if cutoff is None or d < cutoff:
                cost[c_idx][d_idx] = d
    return n, cost


def calculate_cost_by_id(distances, demand, cabs):
    """procedure.py:6-12: fill n*n, cells addressed by the id column."""
    n = len(cabs) if len(cabs) > len(demand) else len(demand)
    cost = [[n * n for _ in range(n)] for _ in range(n)]
    for c_id, _c_frm, c_to in cabs:
        for d_id, d_frm, _d_to in demand:
            cost[c_id][d_id] = int(distances[c_to][d_frm])
    return n, cost


def calculate_cost_np(dist, cab_to, cust_from, fill=BIG_COST, cutoff=None):
    """Vectorised equivalent used for large shapes (checked against calculate_cost in tests)."""
    dist = np.asarray(dist, dtype=np.int32)
    cab_to = np.asarray(cab_to, dtype=np.int64)
    cust_from = np.asarray(cust_from, dtype=np.int64)
    n = max(len(cab_to), len(cust_from))
    cost = np.full((n, n), fill, dtype=np.int32)
    if len(cab_to) and len(cust_from):
        block = dist[cab_to[:, None], cust_from[None, :]]
        if cutoff is not None:
            block = np.where(block < cutoff, block, np.int32(fill))
        cost[: len(cab_to), : len(cust_from)] = block
    return n, cost
