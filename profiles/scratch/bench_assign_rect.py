import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g, cost_ref
eng = td.engine()
rng = np.random.default_rng(5)
dist = g.stand_distances(50)
for n_cabs, n_cust in ((600, 218), (600, 351), (1300, 700), (218, 600), (700, 1300)):
    n, cost = cost_ref.calculate_cost_np(dist, rng.integers(0, 50, n_cabs), rng.integers(0, 50, n_cust), cutoff=10)
    c = torch.from_numpy(np.ascontiguousarray(cost.astype(np.int32))).cuda()
    for mode in ("padded", "rect"):
        kw = {} if mode == "padded" else dict(n_real_rows=n_cabs if n_cabs < n else None, n_real_cols=n_cust if n_cust < n else None)
        ts = []
        for r in range(20):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); col, obj, x, st = eng.assign(c, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        col, obj, x, st = eng.assign(c, want_stats=True, **kw)
        print(f"{n_cabs}x{n_cust} {mode}: ms_med={np.median(ts):.3f} obj={int(obj.item())} phases={st.phases} levels={st.search_steps} free0={st.unassigned_after_auction}", flush=True)
