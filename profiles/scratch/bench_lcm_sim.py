import sys, numpy as np, torch
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g, cost_ref
eng=td.engine()
rng=np.random.default_rng(1)
dist=g.stand_distances(50)
for n_cabs,n_cust in ((1300,142),(877,172),(900,600),(800,836)):
    n,C=cost_ref.calculate_cost_np(dist, rng.integers(0,50,n_cabs), rng.integers(0,50,n_cust), cutoff=10)
    c=torch.from_numpy(C).cuda()
    kw=dict(mask_value=250000, stop_at_value=250000, residual_size=600)
    for _ in range(3): r=eng.lcm(c,**kw)
    torch.cuda.synchronize(); ms=[]
    for _ in range(10):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); r=eng.lcm(c,**kw); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    h=eng.lcm_host_view(*r)
    print(n_cabs,n_cust,'n',n,'median ms',round(float(np.median(ms)),3),'pairs',h['n_pairs'])
