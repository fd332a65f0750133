import sys, numpy as np, torch
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
eng=td.engine()
cases=[('5a-8000',g.config5a(8000),dict(mask_value=100)),('5b-8000',g.config5b_cost(8000),dict(mask_value=250000,sum_below=250000)),
       ('5a-20000',g.config5a(),dict(mask_value=100)),('5b-20000',g.config5b_cost(),dict(mask_value=250000,sum_below=250000))]
for name,C,kw in cases:
    c=torch.from_numpy(C).cuda()
    r=eng.lcm(c,**kw)
    torch.cuda.synchronize()
    ms=[]
    for _ in range(3):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); r=eng.lcm(c,**kw); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    h=eng.lcm_host_view(*r)
    print(name,'median ms',np.median(ms),'total',h['total'],'pairs',h['n_pairs'],flush=True)
    del c
