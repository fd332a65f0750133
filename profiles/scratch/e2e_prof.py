import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import taxidispatcher_b200 as td
from oracle import gen_inputs as g
dem = g.pool_demand(); dist = g.stand_distances(50)
for _ in range(5): td.find_pool_all(dem, dist, 4)
t=time.perf_counter()
for _ in range(200): td.find_pool_all(dem, dist, 4)
print("avg ms", (time.perf_counter()-t)/200*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): td.find_pool_all(dem, dist, 4)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
