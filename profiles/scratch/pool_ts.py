import sys, numpy as np, torch
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
eng=td.engine()
dem=torch.from_numpy(g.pool_demand()).cuda(); dist=torch.from_numpy(g.stand_distances(50)).cuda()
for rep in range(3):
    out,cnt,st=eng.pool_find_shards(dem,dist,4,0,8,8)
torch.cuda.synchronize()
ws=eng._ws['pool']
ts=ws[:256].cpu().numpy().view(np.uint64)
t0=int(ts[31]); seq=[int(x) for x in ts[:31] if x>0]
print('start->first stamp (init+thresholds) us', (seq[0]-t0)/1e3)
names=['partition']
d=[(seq[i+1]-seq[i])/1e3 for i in range(len(seq)-1)]
print('deltas us', [round(x,1) for x in d], 'total', round((seq[-1]-t0)/1e3,1))
print('rounds', st[0].rounds, [s.kept for s in st])
