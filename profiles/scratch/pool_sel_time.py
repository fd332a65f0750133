import sys, numpy as np, torch, ctypes
sys.path.insert(0,'.')
import taxidispatcher_b200 as td
from taxidispatcher_b200 import _lib
from oracle import gen_inputs as g
eng=td.engine(); lib=_lib.lib()
dem=torch.from_numpy(g.pool_demand()).cuda(); dist=torch.from_numpy(g.stand_distances(50)).cuda()
for count in (1,2,8):
    eng.pool_find_shards(dem,dist,4,0,count,8)   # sizes the record list
    for _ in range(3): eng.pool_find_shards(dem,dist,4,0,count,8,want_stats=False)
    torch.cuda.synchronize()
    lib.td_prof_reset(); lib.td_prof_enable(1)
    for _ in range(20): eng.pool_find_shards(dem,dist,4,0,count,8,want_stats=False)
    torch.cuda.synchronize()
    pm,pc=ctypes.c_double(),ctypes.c_int64()
    lib.td_prof_read(_lib.PROF_POOL_ENUM,ctypes.byref(pm),ctypes.byref(pc)); en=pm.value/pc.value
    lib.td_prof_read(_lib.PROF_POOL_SELECT,ctypes.byref(pm),ctypes.byref(pc)); se=pm.value/pc.value
    lib.td_prof_enable(0)
    print(f"shards={count}: enum {en*1e3:.0f} us  select {se*1e3:.0f} us")
