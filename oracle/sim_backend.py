"""CPU backend for taxidispatcher_b200.simulator.Simulator built from the oracles.  TEST INFRASTRUCTURE ONLY:
it lets tests/ check the Simulator restatement's host logic against the reference's golden log
(simulations/simulog_solv.txt, KAT S1) without a GPU, and it is bench.py's CPU baseline for config 4."""
from __future__ import annotations

import time

import numpy as np

from . import assign_ref, cost_ref, lcm_ref, pool_ref

BIG_COST = 250000
DROP_TIME = 10
MAX_NON_LCM = 600


class OracleBackend:
    def __init__(self):
        self.times = {"pool": 0.0, "cost": 0.0, "lcm": 0.0, "solve": 0.0}

    def pool_pairs(self, frm, to, dist):
        t0 = time.perf_counter()
        r = pool_ref.pairs(frm, to, dist, accept_all=True)              # Simulator.java:691
        self.times["pool"] += time.perf_counter() - t0
        return r

    def cost(self, dist, cab_to, cust_from):
        t0 = time.perf_counter()
        n, c = cost_ref.calculate_cost_np(dist, cab_to, cust_from, fill=BIG_COST, cutoff=DROP_TIME)
        self.times["cost"] += time.perf_counter() - t0
        return c

    def lcm_java(self, cost):
        t0 = time.perf_counter()
        o = lcm_ref.lcm_c(cost, BIG_COST, stop_at_value=BIG_COST, residual_size=MAX_NON_LCM)
        self.times["lcm"] += time.perf_counter() - t0
        mn = o["last_min"]
        return list(zip(o["rows"].tolist(), o["cols"].tolist())), (BIG_COST if mn >= BIG_COST else mn)

    def solve(self, n, cost, n_cabs=None, n_cust=None):
        t0 = time.perf_counter()
        if n == 0:
            return []
        _, col = assign_ref.solve_scipy(cost)
        x = assign_ref.x_from_cols(col)
        self.times["solve"] += time.perf_counter() - t0
        return x
