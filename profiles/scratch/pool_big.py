"""Large pool runs on the GPU: timing + size-independent property checks (no CPU reference at this size)."""
import sys, time, json, os
import numpy as np, torch
sys.path.insert(0, '.')
import taxidispatcher_b200 as td
from oracle import gen_inputs as g, pool_ref

def check_properties(dem, dist, plans_by_shard, k=4, n_shards=8, residual_check=True):
    n = len(dem)
    F, T, W, L = dem[:, 1], dem[:, 2], dem[:, 3], dem[:, 4]
    step = n // n_shards + 1
    for sh, plans in enumerate(plans_by_shard):
        if len(plans) == 0: continue
        p = plans[:, :k]; q = plans[:, k:2*k]; cost = plans[:, 8]
        assert (np.diff(cost) >= 0).all(), "not sorted by cost"
        assert len(set(p.ravel().tolist())) == p.size, "customers repeated inside a shard"
        assert ((p[:, 0] >= step*sh) & (p[:, 0] < min(n, step*(sh+1)))).all(), "leader outside the shard"
        assert (np.sort(p, 1) == np.sort(q, 1)).all()
        # feasibility of every kept plan (wait rule + detour rule), formulas of SURVEY 8(a)
        D = dist
        legs = D[F[p[:, :-1]], F[p[:, 1:]]]
        cum = np.concatenate([np.zeros((len(p), 1), int), np.cumsum(legs, 1)], 1)
        assert (cum <= W[p]).all(), "wait rule violated"
        first = D[F[p[:, -1]], T[q[:, 0]]]
        drops = D[T[q[:, :-1]], T[q[:, 1:]]]
        total = legs.sum(1) + first + drops.sum(1)
        assert (total == cost).all(), "cost mismatch"
        dcum = np.concatenate([np.zeros((len(p), 1), int), np.cumsum(drops, 1)], 1) + first[:, None]
        for d in range(k):
            c = q[:, d]
            pos = (p == c[:, None]).argmax(1)
            suffix = np.array([legs[i, pos[i]:].sum() for i in range(len(p))])
            ride = suffix + dcum[:, d]
            lim = D[F[c], T[c]] * (1 + L[c] / 100.0)
            assert (ride <= lim).all(), "detour rule violated"
        if residual_check:
            used = np.zeros(n, bool); used[p.ravel()] = True
            rest = np.nonzero(~used)[0]
            sub = dem[rest].copy()
            # any feasible plan among the unused customers whose leader lies in this shard would contradict maximality
            if len(rest) <= 400:
                plans_r, _ = pool_ref.find(sub, dist, k, 0, 1, dedup=False, cap=1 << 20)
                lead = rest[plans_r[:, 0]] if len(plans_r) else np.array([], int)
                assert not ((lead >= step*sh) & (lead < min(n, step*(sh+1)))).any(), "greedy scan is not maximal"
    return True

def run(n, mf, seed=None, golden=None):
    seed = n if seed is None else seed
    dem = g.pool_demand(n, seed=seed); dist = g.stand_distances(50)
    eng = td.engine()
    dd = torch.from_numpy(dem).cuda(); ds = torch.from_numpy(dist).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, cnt, st = eng.pool_find_shards(dd, ds, 4, 0, 8, 8, max_feasible=mf)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    counts = cnt.cpu().numpy(); plans = [out[s, :counts[s]].cpu().numpy() for s in range(8)]
    ev = sum(s.evaluated for s in st); fe = sum(s.feasible for s in st)
    print(f"n={n} mf={mf}: {dt*1e3:.1f} ms  evaluated={ev:.4g} ({ev/dt:.3g} plans/s) feasible={fe:.4g} kept={[int(c) for c in counts]} passes={st[0].passes} rounds={st[0].rounds}", flush=True)
    if golden:
        gold = json.load(open(golden))
        for s in range(8):
            assert plans[s].tolist() == gold["shards"][s]["plans"], s
            ge = gold["shards"][s]["stats"]["evaluated"]   # the reference's int32 counter wraps (pool_n.c:28)
            assert ge is None or (st[s].evaluated - ge) % (1 << 32) == 0, (st[s].evaluated, ge)
            assert st[s].feasible == gold["shards"][s]["stats"]["feasible"]
        print("   matches golden", golden)
    t0 = time.perf_counter(); check_properties(dem, dist, plans); print(f"   properties ok ({time.perf_counter()-t0:.1f}s)", flush=True)

for a in sys.argv[1:]:
    n, mf = a.split(':')[:2]
    gold = a.split(':')[2] if a.count(':') >= 2 else None
    run(int(n), int(float(mf)), golden=gold)
