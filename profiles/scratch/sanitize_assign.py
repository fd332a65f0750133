import sys
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from oracle import assign_ref
rng = np.random.default_rng(3)
for n in (4, 8, 64, 132, 260, 301, 512):
    for C in (rng.integers(1, 40, (n, n)), rng.integers(0, 3, (n, n)), np.abs(rng.integers(0, 50, n)[:, None] - rng.integers(0, 50, n)[None, :])):
        C = C.astype(np.int32)
        x, col, obj, st = td.solve_full(n, C)
        ref = assign_ref.solve_scipy(C)[0]
        assert obj == ref, (n, obj, ref)
print("ok")
