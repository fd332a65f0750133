#!/usr/bin/env python
"""Phase timing of pool_select from the %globaltimer stamps the kernel leaves in its control block, and CUDA-event timing
of the enumeration / selection launches (td_prof), for the 8-shard call (N = 1) and a single-shard call (N = 8).
Run on the GPU box:  python profiles/pool_ts.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taxidispatcher_b200 as td  # noqa: E402
from taxidispatcher_b200 import _lib  # noqa: E402
from oracle import gen_inputs as g  # noqa: E402

eng = td.engine()
lib = _lib.lib()
dem = torch.from_numpy(g.pool_demand()).cuda()
dist = torch.from_numpy(g.stand_distances(50)).cuda()
for shards in ((0, 8), (0, 1)):
    for rep in range(3):
        out, cnt, st = eng.pool_find_shards(dem, dist, 4, shards[0], shards[1], 8)
    lib.td_prof_reset(); lib.td_prof_enable(1)
    for rep in range(10):
        out, cnt, tok = eng.pool_find_shards(dem, dist, 4, shards[0], shards[1], 8, defer_stats=True)
    torch.cuda.synchronize()
    pm, pc = ctypes.c_double(), ctypes.c_int64()
    for name, kind in (("enum", _lib.PROF_POOL_ENUM), ("select", _lib.PROF_POOL_SELECT)):
        lib.td_prof_read(kind, ctypes.byref(pm), ctypes.byref(pc))
        print("shards %s: %s avg %.1f us over %d launches" % (shards, name, 1e3 * pm.value / max(pc.value, 1), pc.value))
    lib.td_prof_enable(0)
    ws = eng._ws["pool"]
    ts = ws[:256].cpu().numpy().view(np.uint64)
    t0 = int(ts[31]); seq = [int(x) for x in ts[:31] if x > 0]
    d = [(seq[i + 1] - seq[i]) / 1e3 for i in range(len(seq) - 1)]
    print("  select phases us: init %.1f | partition, then (filter, rounds) per band:" % ((seq[0] - t0) / 1e3), [round(x, 1) for x in d],
          "total", round((seq[-1] - t0) / 1e3, 1))
    print("  rounds", st[0].rounds, "kept", [int(s.kept) for s in st])
