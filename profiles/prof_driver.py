#!/usr/bin/env python
"""Small, deterministic driver for ncu captures: one invocation of each hot-path kernel on its
BASELINE.json shape.  Usage (on the GPU box, see profiles/README.md):
    python profiles/prof_driver.py [pool] [pool8] [cost] [lcm] [assign5a] [assign5b] [assign2s]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import taxidispatcher_b200 as td  # noqa: E402
from oracle import gen_inputs as g  # noqa: E402

which = sys.argv[1:] or ["pool", "cost", "lcm", "assign5a"]
eng = td.engine()
if "pool" in which:
    dem = torch.from_numpy(g.pool_demand()).cuda()
    dist = torch.from_numpy(g.stand_distances(50)).cuda()
    for sh in (0, 1):
        out, cnt, st = eng.pool_find(dem, dist, 4, sh, 8)
        print("pool shard", sh, st.evaluated, st.feasible, st.kept, st.rounds)
if "pool8" in which:   # the launch bench.py times at N = 1: all 8 logical shards in one enumeration + one selection
    dem = torch.from_numpy(g.pool_demand()).cuda()
    dist = torch.from_numpy(g.stand_distances(50)).cuda()
    for _ in range(2):   # the first call sizes the record list (cost windows), the second one runs single-pass
        out, cnt, st = eng.pool_find_shards(dem, dist, 4, 0, 8, 8)
    print("pool8", sum(s.evaluated for s in st), sum(s.feasible for s in st), [int(s.kept) for s in st], "passes", st[0].passes)
    # the asynchronous single pass bench.py times: the LAST pool_enum / pool_select launches of the capture
    out, cnt, tok = eng.pool_find_shards(dem, dist, 4, 0, 8, 8, defer_stats=True)
    st2, ov = eng.pool_read_stats(tok)
    print("pool8 async", sum(s.evaluated for s in st2), "overflow", ov)
if "pool5k" in which:   # north-star size: one 512-way slice (10 leading customers) of the 5000-customer input, single pass
    dem = torch.from_numpy(g.pool_demand(5000, seed=5000)).cuda()
    dist = torch.from_numpy(g.stand_distances(50)).cuda()
    out, cnt, st = eng.pool_find_shards(dem, dist, 4, 255, 1, 512, max_feasible=120_000_000)
    print("pool5k slice 255", st[0].evaluated, st[0].feasible, st[0].kept, "passes", st[0].passes)
if "cost" in which:
    cab_to, cust_from = g.config5b()
    d = torch.from_numpy(g.stand_distances(4000)).cuda()
    out = None
    for _ in range(3):
        out = eng.cost_matrix(d, torch.from_numpy(cab_to).cuda(), torch.from_numpy(cust_from).cuda(), out=out)
    torch.cuda.synchronize()
    print("cost", int(out[123, 456]))
if "lcm" in which:
    c = torch.from_numpy(g.config2()).cuda()
    for _ in range(2):
        r = eng.lcm_host_view(*eng.lcm(c, 100))
    print("lcm", r["total"])
if "assign5a" in which:
    c = torch.from_numpy(g.config5a()).cuda()
    col, obj, _, st = eng.assign(c, want_stats=True)
    print("assign5a", int(obj.item()), st.phases, st.search_steps, st.rows_scanned)
if "assign5b" in which:
    c = torch.from_numpy(g.config5b_cost()).cuda()
    col, obj, _, st = eng.assign(c, want_stats=True)
    print("assign5b", int(obj.item()), st.phases, st.search_steps, st.rows_scanned)
if "assign2s" in which:
    c = torch.from_numpy(g.config2_stand()).cuda()
    col, obj, _, st = eng.assign(c, want_stats=True)
    print("assign2s", int(obj.item()), st.phases, st.search_steps, st.rows_scanned)
torch.cuda.synchronize()
