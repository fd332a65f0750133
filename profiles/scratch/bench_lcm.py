import sys, numpy as np, torch
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
eng=td.engine()
for name,C,kw in (('cfg2',g.config2(),dict(mask_value=100)),('cfg2stand',g.config2_stand(),dict(mask_value=250000,sum_below=250000))):
    c=torch.from_numpy(C).cuda()
    for _ in range(3): r=eng.lcm(c,**kw)
    torch.cuda.synchronize()
    ms=[]
    for _ in range(10):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); r=eng.lcm(c,**kw); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    h=eng.lcm_host_view(*r)
    print(name,'median ms',np.median(ms),'total',h['total'],'pairs',h['n_pairs'])
