"""Where the 5000-customer job spends its time: CUDA-event totals of the enumeration / selection launches (td_prof)
for two record capacities.  Run on the GPU box:  python profiles/scratch/pool5k_split.py"""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import taxidispatcher_b200 as td
from taxidispatcher_b200 import _lib
from oracle import gen_inputs as g
eng = td.engine(); lib = _lib.lib()
n = 5000
dem = torch.from_numpy(g.pool_demand(n, seed=n)).cuda(); dist = torch.from_numpy(g.stand_distances(50)).cuda()
for records in [int(x) for x in sys.argv[1:]] or [1_600_000_000]:
    try:
        eng._ws.pop("pool", None); torch.cuda.empty_cache()
        eng._workspace("pool", eng.lib.td_pool_shards_workspace_bytes(n, 50, 4, 8, records))
    except torch.OutOfMemoryError:
        print(records, "OOM"); continue
    torch.cuda.synchronize()
    lib.td_prof_reset(); lib.td_prof_enable(1)
    t = time.perf_counter()
    out, cnt, st = eng.pool_find_shards(dem, dist, 4, 0, 8, 8, max_feasible=records)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t
    pm, pc = ctypes.c_double(), ctypes.c_int64()
    res = {}
    for name, kind in (("enum", _lib.PROF_POOL_ENUM), ("select", _lib.PROF_POOL_SELECT)):
        lib.td_prof_read(kind, ctypes.byref(pm), ctypes.byref(pc)); res[name] = (round(pm.value, 1), pc.value)
    lib.td_prof_enable(0)
    print(records, "sec %.3f" % sec, "passes", st[0].passes, res, "kept", int(cnt.sum()))
