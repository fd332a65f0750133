import sys, numpy as np, torch
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
eng=td.engine()
dem=torch.from_numpy(g.pool_demand()).cuda(); dist=torch.from_numpy(g.stand_distances(50)).cuda()
out,cnt,st=eng.pool_find_shards(dem,dist,4,0,8,8)
for _ in range(3): eng.pool_merge_padded(out,cnt,None,722,4)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): m,c=eng.pool_merge_padded(out,cnt,None,722,4)
b.record(); torch.cuda.synchronize()
print("merge us", a.elapsed_time(b)/50*1e3, int(c.item()))
