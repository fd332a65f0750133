import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g, cost_ref
eng = td.engine()
rng = np.random.default_rng(5)
def run(name, C, reps=20):
    c = torch.from_numpy(np.ascontiguousarray(C.astype(np.int32))).cuda()
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        col, obj, x, st = eng.assign(c, want_stats=False)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    col, obj, x, st = eng.assign(c, want_stats=True)
    print(f"{name} n={C.shape[0]} ms_med={np.median(ts):.3f} min={min(ts):.3f} obj={int(obj.item())} phases={st.phases} levels={st.search_steps} free0={st.unassigned_after_auction}", flush=True)
dist = g.stand_distances(50)
for n_cabs, n_cust in ((600, 218), (600, 351), (1300, 700), (200, 200), (64, 64)):
    n, cost = cost_ref.calculate_cost_np(dist, rng.integers(0, 50, n_cabs), rng.integers(0, 50, n_cust), cutoff=10)
    run(f"sim {n_cabs}x{n_cust}", cost)
run("uniform 600", rng.integers(1, 40, (600, 600)))
run("uniform 2000", g.config2())
