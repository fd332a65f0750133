import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
eng = td.engine()
def run(name, C, reps=3):
    c = torch.from_numpy(np.ascontiguousarray(C)).cuda()
    torch.cuda.synchronize()
    for r in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        col, obj, x, st = eng.assign(c, want_stats=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        n = C.shape[0]
        print(f"{name} n={n} ms={ms:.2f} obj={int(obj.item())} phases={st.phases} levels={st.search_steps} rows={st.rows_scanned} ({st.rows_scanned/n:.1f} sweeps, {4*n*st.rows_scanned/ms/1e6:.0f} GB/s) free0={st.unassigned_after_auction} aug={st.augmentations}", flush=True)
which = sys.argv[1:] or ['2', '2s', '5a5', '5b5', '5a', '5b']
for w in which:
    if w == '1a': run('1a', g.config1a())
    if w == '2': run('cfg2', g.config2())
    if w == '2s': run('cfg2stand', g.config2_stand())
    if w == '5a5': run('5a-5000', g.config5a(5000))
    if w == '5b5': run('5b-5000', g.config5b_cost(5000))
    if w == '5a': run('5a-20000', g.config5a())
    if w == '5b': run('5b-20000', g.config5b_cost())
