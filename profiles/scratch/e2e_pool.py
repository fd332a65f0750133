import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from taxidispatcher_b200 import dispatch
from oracle import gen_inputs as g
dem = g.pool_demand(722); dist = g.stand_distances(50)
eng = td.engine()
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plans, st = dispatch.find_pool_all(dem, dist, 4)
    dt = time.perf_counter() - t0
    print(i, f"{dt*1e3:.2f} ms", len(plans), st.get("rounds"), {k: v for k, v in eng._ws.items() if isinstance(k, tuple)}, flush=True)
