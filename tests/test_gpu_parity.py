"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors.  Bit-exact bar for every integer result."""
import hashlib

import numpy as np
import pytest

from oracle import assign_ref, cost_ref, gen_inputs as g, lcm_ref, pool_ref
from conftest import load_golden

pytestmark = pytest.mark.gpu


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(arr, dtype=np.int32)).tobytes()).hexdigest()


# ---- K1 --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_cabs,n_cust,S,cutoff", [(3, 4, 10, None), (190, 200, 50, None), (200, 190, 50, 10),
                                                    (1, 1, 2, None), (257, 131, 4000, None), (131, 258, 20, 5),
                                                    (1300, 722, 50, 10), (33, 0, 5, None), (0, 33, 5, None)])
def test_cost_matrix_matches_reference_loops(td, n_cabs, n_cust, S, cutoff):
    rng = np.random.default_rng(n_cabs * 1000 + n_cust)
    dist = g.stand_distances(S)
    cabs = [(i, int(rng.integers(0, S)), int(rng.integers(0, S))) for i in range(n_cabs)]
    dem = [(i, int(rng.integers(0, S)), int(rng.integers(0, S))) for i in range(n_cust)]
    n, cost = td.calculate_cost(dist, dem, cabs, cutoff=cutoff)
    n_ref, c_ref = cost_ref.calculate_cost(dist, dem, cabs, cutoff=cutoff)
    assert n == n_ref
    assert np.array_equal(np.asarray(cost), np.asarray(c_ref, dtype=np.int32))


def test_cost_matrix_kat_a3_and_asymmetric_table(td):
    d10 = g.stand_distances(10)
    n, cost = td.calculate_cost(d10, [(0, 0, 2), (1, 0, 5), (2, 3, 1), (3, 5, 1)], [(0, 3, 3), (1, 3, 1), (2, 0, 5)], fill=16)
    assert cost.tolist() == [[3, 3, 0, 2], [1, 1, 2, 4], [5, 5, 2, 0], [16, 16, 16, 16]]
    rng = np.random.default_rng(9)
    dist = rng.integers(0, 1000, (37, 37)).astype(np.int32)  # not symmetric: catches a transposed lookup
    cabs = [(i, 0, int(rng.integers(0, 37))) for i in range(45)]
    dem = [(i, int(rng.integers(0, 37)), 0) for i in range(51)]
    n, cost = td.calculate_cost(dist, dem, cabs)
    assert np.array_equal(cost, np.asarray(cost_ref.calculate_cost(dist, dem, cabs)[1], dtype=np.int32))
    assert td.calculate_cost(dist, [], []) == (0, 0)
    with pytest.raises(IndexError):
        td.calculate_cost(dist, [(0, 99, 0)], [(0, 0, 1)])


def test_cost_matrix_full_size_properties(td):
    """config 5-B shape (20k x 20k, 4000 stands): too big for the Python loops, so check against the
    vectorised oracle on sampled rows plus a checksum of the whole matrix."""
    import torch
    cab_to, cust_from = g.config5b()
    dist = g.stand_distances(4000)
    eng = td.engine()
    out = eng.cost_matrix(torch.from_numpy(dist).cuda(), torch.from_numpy(cab_to).cuda(), torch.from_numpy(cust_from).cuda())
    total = int(out.sum(dtype=torch.int64).item())
    ref_total = int(np.abs(cab_to[:, None].astype(np.int64) - cust_from[None, :]).sum())
    assert total == ref_total
    rows = [0, 1, 7777, 19999]
    got = out[rows].cpu().numpy()
    assert np.array_equal(got, np.abs(cab_to[rows][:, None] - cust_from[None, :]))


def test_cost_matrix_row_blocks(td):
    """Row blocks (the multi-GPU split of split.py:129-134) concatenate to the whole matrix: dummy rows, dummy columns,
    the cutoff, an asymmetric table, n % 4 != 0 (scalar head / tail of every row) and blocks that start inside the
    dummy rows."""
    import torch
    from taxidispatcher_b200 import parallel as P
    eng = td.engine()
    rng = np.random.default_rng(404)
    for n_cabs, n_cust, S, cutoff in ((5000, 4600, 300, None), (4096, 5000, 1000, 40), (4999, 4999, 64, None), (700, 650, 50, 10)):
        dist = rng.integers(0, 90, (S, S)).astype(np.int32)
        cab_to = rng.integers(0, S, n_cabs).astype(np.int32)
        cust_from = rng.integers(0, S, n_cust).astype(np.int32)
        n = max(n_cabs, n_cust)
        ref = np.full((n, n), g.BIG_COST, np.int32)
        blk = dist[cab_to][:, cust_from]
        if cutoff is not None:
            blk = np.where(blk < cutoff, blk, g.BIG_COST)
        ref[:n_cabs, :n_cust] = blk
        dd, dc, du = torch.from_numpy(dist).cuda(), torch.from_numpy(cab_to).cuda(), torch.from_numpy(cust_from).cuda()
        whole = eng.cost_matrix(dd, dc, du, cutoff=cutoff).cpu().numpy()
        assert np.array_equal(whole, ref), (n_cabs, n_cust, S)
        for w in (2, 3):
            parts = [eng.cost_matrix(dd, dc, du, cutoff=cutoff, rows=P.rows_for_rank(n, r, w)).cpu().numpy() for r in range(w)]
            assert np.array_equal(np.concatenate(parts, axis=0), ref), (n_cabs, n_cust, S, w)
    # the north-star shape through the public sharded entry point (one rank here: the whole matrix)
    cab_to, cust_from = g.config5b()
    n, (lo, hi), block = P.cost_matrix_sharded(g.stand_distances(4000), cab_to, cust_from)
    assert (n, lo, hi) == (20000, 0, 20000)
    assert int(block.sum(dtype=torch.int64).item()) == int(np.abs(cab_to[:, None].astype(np.int64) - cust_from[None, :]).sum())
    rows = [0, 5, 12345, 19999]
    assert np.array_equal(block[rows].cpu().numpy(), np.abs(cab_to[rows][:, None] - cust_from[None, :]))


# ---- K3 --------------------------------------------------------------------------------------------
def _check_lcm(td, cost, **kw):
    n = cost.shape[0]
    r = td.LCM(n, cost, **kw)
    o = lcm_ref.lcm_c(cost, kw.get("mask_value", g.BIG_COST), kw.get("stop_above", 2**31 - 1),
                      kw.get("stop_at_value", 2**31 - 1), kw.get("sum_below", 2**31 - 1), kw.get("residual_size", 0),
                      kw.get("max_iters", -1))
    assert r["n_pairs"] == o["n_pairs"], (r["n_pairs"], o["n_pairs"])
    assert np.array_equal(r["rows"], o["rows"]) and np.array_equal(r["cols"], o["cols"])
    assert r["total"] == o["total"]
    assert r["last_min"] == o["last_min"]
    return r


@pytest.mark.parametrize("n", [1, 2, 3, 31, 100, 257, 600])
def test_lcm_heuristic_variant(td, n):
    c = np.random.default_rng(n).integers(1, 40, (n, n)).astype(np.int32)
    _check_lcm(td, c, mask_value=100)


def test_lcm_all_equal_and_ties(td):
    for n in (1, 5, 64, 130):
        _check_lcm(td, np.full((n, n), 7, np.int32), mask_value=100)
        _check_lcm(td, np.zeros((n, n), np.int32), mask_value=g.BIG_COST, sum_below=g.BIG_COST)
    c = np.random.default_rng(3).integers(0, 3, (90, 90)).astype(np.int32)
    _check_lcm(td, c, mask_value=100)


def test_lcm_mask_is_a_value_tail(td):
    """dummy rows of big_cost: once every free cell equals the mask the argmin re-selects masked cells
    (SURVEY section 4 trap 4); also masks below / equal to live costs."""
    dist, cab_to, cust_from = g.config1b()
    n, cost = cost_ref.calculate_cost_np(dist, cab_to, cust_from)
    r = _check_lcm(td, cost, mask_value=g.BIG_COST, sum_below=g.BIG_COST)          # split.py:161-175
    gold = load_golden("lcm.json")["config1b_split"]
    assert r["total"] == gold["total"] and r["rows"].tolist() == gold["rows"] and r["cols"].tolist() == gold["cols"]
    n2, cost2 = cost_ref.calculate_cost_np(dist, cab_to[:150], cust_from)           # 50 dummy rows
    _check_lcm(td, cost2, mask_value=g.BIG_COST, sum_below=g.BIG_COST)
    n3, cost3 = cost_ref.calculate_cost_np(dist, cab_to, cust_from[:120], cutoff=10)  # dummy columns + cutoff holes
    _check_lcm(td, cost3, mask_value=g.BIG_COST, sum_below=g.BIG_COST)
    c = np.random.default_rng(8).integers(1, 40, (80, 80)).astype(np.int32)
    for mask in (20, 1, 0, 39, 40):   # mask inside / below the live value range
        _check_lcm(td, c, mask_value=mask)
    c[0, :] = 50                       # row 0 expensive: exercises the (0, min masked column) branch
    _check_lcm(td, c, mask_value=45)
    _check_lcm(td, c, mask_value=45, max_iters=60)


def test_lcm_threshold_and_java_variants(td):
    dist, cab_to, cust_from = g.config1b()
    n, cost = cost_ref.calculate_cost_np(dist, cab_to, cust_from, cutoff=10)
    for thr in (0, 3, 10, 20):
        _check_lcm(td, cost, mask_value=g.BIG_COST, stop_above=thr, sum_below=g.BIG_COST)  # greedy_opt.py / simulate.py
    for res in (0, 150, 199, 1, 500):
        _check_lcm(td, cost, mask_value=g.BIG_COST, stop_at_value=g.BIG_COST, residual_size=res)  # Simulator.java
    pairs, mn = td.LCM_java(cost, max_non_lcm=150)
    gp, gm = lcm_ref.lcm_java(cost, max_non_lcm=150)
    assert pairs == gp and mn == gm
    gj = load_golden("lcm.json")["config1b_java_res150"]
    n0, cost0 = cost_ref.calculate_cost_np(dist, cab_to, cust_from)
    pairs, mn = td.LCM_java(cost0, max_non_lcm=150)
    assert [list(p) for p in pairs] == gj["pairs"] and mn == gj["lcm_min_val"]
    tot, sup, dem = td.LCM_greedy_opt(n, cost, threshold=3)
    assert (tot, sup, dem) == tuple(int(v) if i == 0 else v for i, v in enumerate(lcm_ref.lcm_greedy_opt(n, cost, 3)))
    assert td.LCM_split(n, cost) == int(lcm_ref.lcm_split(n, cost)[0])
    assert td.LCM_heuristic(50, cost[:50, :50].copy(), mask=100) == int(lcm_ref.lcm_heuristic(50, cost[:50, :50].reshape(-1))[0])


def test_lcm_config2_golden_trace(td):
    gold = load_golden("lcm.json")
    r = td.LCM(2000, g.config2(), mask_value=100)
    gh = gold["config2_heuristic"]
    assert r["total"] == gh["total"] == 2111
    assert sha(r["rows"]) == gh["rows_sha256"] and sha(r["cols"]) == gh["cols_sha256"]
    Cs = g.config2_stand()
    r = td.LCM(2000, Cs, mask_value=g.BIG_COST, sum_below=g.BIG_COST)
    gs = gold["config2_stand_split"]
    assert r["total"] == gs["total"] and sha(r["rows"]) == gs["rows_sha256"] and sha(r["cols"]) == gs["cols_sha256"]
    r = td.LCM(2000, Cs, mask_value=g.BIG_COST, stop_above=10, sum_below=g.BIG_COST)
    gt = gold["config2_stand_greedy_opt_thr10"]
    assert (r["total"], r["n_pairs"]) == (gt["total"], gt["n_pairs"]) and sha(r["rows"]) == gt["rows_sha256"]


# ---- K4 --------------------------------------------------------------------------------------------
def test_pool_small_golden_all_shards(td):
    for case in load_golden("pool_small.json"):
        dem = np.array(case["demand"], dtype=np.int32)
        dist = g.stand_distances(case["n_stands"])
        k = case["pool_size"]
        shard_plans = []
        for s in case["shards"]:
            plans, st = td.find_pool(dem, dist, k, s["shard"])
            assert {q: st[q] for q in ("evaluated", "feasible", "kept")} == s["stats"], (case["label"], s["shard"])
            assert plans.tolist() == s["plans"], (case["label"], s["shard"])
            shard_plans.append(plans)
        assert td.pool_merge(shard_plans, len(dem), k).tolist() == case["merged"], case["label"]
        merged, stats = td.find_pool_all(dem, dist, k)
        assert merged.tolist() == case["merged"]


def test_pool_config3_722_golden(td):
    gold = load_golden("pool722.json")
    dem = g.pool_demand()
    dist = g.stand_distances(50)
    shard_plans = []
    for s in gold["shards"]:
        plans, st = td.find_pool(dem, dist, 4, s["shard"])
        assert st["evaluated"] == s["stats"]["evaluated"] and st["feasible"] == s["stats"]["feasible"]
        assert plans.tolist() == s["plans"], s["shard"]
        shard_plans.append(plans)
    merged = td.pool_merge(shard_plans, 722, 4)
    assert merged.tolist() == gold["merged"]
    assert len(merged) == 110 and int(merged[:, 8].sum()) == 1840


@pytest.mark.parametrize("k", [2, 3, 4])
def test_pool_random_tables_vs_oracle(td, k):
    """asymmetric random distance table, ragged waits / losses, several shard counts"""
    rng = np.random.default_rng(40 + k)
    S = 23
    dist = rng.integers(0, 9, (S, S)).astype(np.int32)
    np.fill_diagonal(dist, 0)
    n = 70
    dem = np.stack([np.arange(n), rng.integers(0, S, n), rng.integers(0, S, n), rng.integers(-1, 12, n),
                    rng.integers(0, 120, n)], axis=1).astype(np.int32)
    for n_shards, shard in ((8, 0), (8, 7), (3, 1), (1, 0)):
        plans, st = td.find_pool(dem, dist, k, shard, n_shards)
        oplans, ost = pool_ref.find(dem, dist, k, shard, n_shards)
        assert {q: st[q] for q in ost} == ost
        assert np.array_equal(plans, oplans)


def test_pool_k4_pruning_bounds_fuzz(td):
    """the exact pruning bounds of the K = 4 enumeration (shortest-path closure, last-dropped passenger) on tables that
    break the triangle inequality in different ways: sparse zeros, one-way shortcuts, constant rows, from == to trips,
    waits at and around the feasibility edge.  Counters and plans must equal the oracle case by case."""
    rng = np.random.default_rng(20261018)
    for case in range(36):
        S = int(rng.integers(3, 30))
        n = int(rng.integers(8, 56))
        kind = case % 4
        if kind == 0:                                    # metric line with random one-way shortcuts
            dist = np.abs(np.arange(S)[:, None] - np.arange(S)[None, :]).astype(np.int32)
            for _ in range(S):
                dist[rng.integers(0, S), rng.integers(0, S)] = rng.integers(0, 3)
        elif kind == 1:                                  # mostly zeros
            dist = (rng.integers(0, 10, (S, S)) * (rng.random((S, S)) < 0.3)).astype(np.int32)
        elif kind == 2:                                  # wide range, asymmetric
            dist = rng.integers(0, 200, (S, S)).astype(np.int32)
        else:                                            # constant rows (every leg out of a stand costs the same)
            dist = np.repeat(rng.integers(1, 12, (S, 1)), S, axis=1).astype(np.int32)
        np.fill_diagonal(dist, 0)
        fr = rng.integers(0, S, n)
        to = np.where(rng.random(n) < 0.1, fr, rng.integers(0, S, n))     # some trips end where they start
        scale = max(1, int(dist.max()))
        dem = np.stack([np.arange(n), fr, to, rng.integers(0, 3 * scale + 1, n), rng.integers(0, 150, n)], axis=1).astype(np.int32)
        n_shards = int(rng.integers(1, 9))
        shard = int(rng.integers(0, n_shards))
        plans, st = td.find_pool(dem, dist, 4, shard, n_shards)
        oplans, ost = pool_ref.find(dem, dist, 4, shard, n_shards)
        assert {q: st[q] for q in ost} == ost, (case, kind, S, n)
        assert np.array_equal(plans, oplans), (case, kind, S, n)


@pytest.mark.parametrize("S", [64, 65, 100, 128, 129])
def test_pool_k4_closure_tile_sizes(td, S):
    """the shortest-path closure runs in 4 x 4 register tiles up to 64 stands and 8 x 8 up to 128; above that the bound
    is switched off.  Non-metric tables with cheap detours through a few hub stands make D* differ from D."""
    rng = np.random.default_rng(1000 + S)
    dist = rng.integers(4, 40, (S, S)).astype(np.int32)
    hubs = rng.choice(S, 5, replace=False)
    dist[hubs, :] = rng.integers(0, 3, (5, S))
    dist[:, hubs] = rng.integers(0, 3, (S, 5))
    np.fill_diagonal(dist, 0)
    n = 44
    dem = np.stack([np.arange(n), rng.integers(0, S, n), rng.integers(0, S, n), rng.integers(0, 30, n),
                    rng.integers(0, 200, n)], axis=1).astype(np.int32)
    plans, st = td.find_pool(dem, dist, 4, 0, 1)
    oplans, ost = pool_ref.find(dem, dist, 4, 0, 1)
    assert {q: st[q] for q in ost} == ost
    assert np.array_equal(plans, oplans)
    assert ost["feasible"] > 0


@pytest.mark.parametrize("k", [2, 3, 4])
def test_pool_large_costs_take_the_two_level_keys(td, k):
    """plan costs >= 255 fall into the open-ended histogram bucket: the selection cannot pack (cost, rank) into one
    64-bit key and uses the two-level minima; waits are long enough for the explicit slack test as well"""
    rng = np.random.default_rng(90 + k)
    S = 17
    dist = rng.integers(40, 160, (S, S)).astype(np.int32)
    np.fill_diagonal(dist, 0)
    n = 60
    dem = np.stack([np.arange(n), rng.integers(0, S, n), rng.integers(0, S, n), rng.integers(100, 400, n),
                    rng.integers(50, 400, n)], axis=1).astype(np.int32)
    for n_shards, shard in ((8, 2), (2, 1), (1, 0)):
        plans, st = td.find_pool(dem, dist, k, shard, n_shards)
        oplans, ost = pool_ref.find(dem, dist, k, shard, n_shards)
        assert {q: st[q] for q in ost} == ost
        assert np.array_equal(plans, oplans)


def test_pool_k4_fixed_point_limits(td):
    """K = 4 evaluates in a x32 fixed point: distances up to 2^22 stay exact, larger ones are refused (not clamped)"""
    from taxidispatcher_b200 import TaxiDispatchError
    rng = np.random.default_rng(7)
    S, n = 9, 26
    dist = rng.integers(1, 1 << 22, (S, S)).astype(np.int32)
    dist[rng.integers(0, S, 6), rng.integers(0, S, 6)] = 1 << 22          # exactly the limit
    np.fill_diagonal(dist, 0)
    dem = np.stack([np.arange(n), rng.integers(0, S, n), rng.integers(0, S, n), rng.integers(1 << 20, 1 << 24, n),
                    rng.integers(0, 300, n)], axis=1).astype(np.int32)
    for k in (3, 4):
        plans, st = td.find_pool(dem, dist, k, 0, 1)
        oplans, ost = pool_ref.find(dem, dist, k, 0, 1)
        assert {q: st[q] for q in ost} == ost and np.array_equal(plans, oplans)
    dist[2, 5] = (1 << 22) + 1
    for k in (2, 3, 4):                                                    # K = 2, 3: the (cost << 5 | perm) key has the same limit
        with pytest.raises(TaxiDispatchError):
            td.find_pool(dem, dist, k, 0, 1)
    dist[2, 5] = -1                                                        # a negative cost would mark a hole in the record list
    with pytest.raises(TaxiDispatchError):
        td.find_pool(dem, dist, 4, 0, 1)
    dist[2, 5] = 3
    dem[7, 1] = S                                                          # stand index outside the table (pool_n.c would read out of bounds)
    with pytest.raises(TaxiDispatchError):
        td.find_pool(dem, dist, 4, 0, 1)
    dem[7, 1] = 0
    dem[3, 2] = -1
    with pytest.raises(TaxiDispatchError):
        td.find_pool(dem, dist, 3, 0, 1)


def test_pool_edge_cases(td):
    dist = g.stand_distances(51)
    plans, st = td.find_pool(np.zeros((0, 5), np.int32), dist, 4, 0)
    assert len(plans) == 0 and st["evaluated"] == 0
    dem = g.pool_demand(3, seed=2)
    for k in (2, 3, 4):
        plans, st = td.find_pool(dem, dist, k, 0)
        oplans, ost = pool_ref.find(dem, dist, k, 0)
        assert np.array_equal(plans, oplans) and st["evaluated"] == ost["evaluated"]
    dem = g.pool_demand(40, seed=2)
    dem[:, 3] = 100   # waits beyond the 64-entry budget table: the explicit slack test path
    dem[:, 4] = 300
    plans, st = td.find_pool(dem, dist, 3, 0, 2)
    oplans, ost = pool_ref.find(dem, dist, 3, 0, 2)
    assert {q: st[q] for q in ost} == ost and np.array_equal(plans, oplans)


# ---- K2 --------------------------------------------------------------------------------------------
def _check_assign(td, cost, ref_obj=None):
    cost = np.asarray(cost, dtype=np.int32)
    n = cost.shape[0]
    x, col, obj, st = td.solve_full(n, cost)
    if ref_obj is None:
        ref_obj = assign_ref.solve_scipy(cost)[0]
    assert obj == ref_obj, (obj, ref_obj)
    assert assign_ref.check_x(x, n)                              # row sums = column sums = 1, binary
    assert sorted(col.tolist()) == list(range(n))
    assert int(cost[np.arange(n), col].sum()) == obj
    assert np.array_equal(np.nonzero(x.reshape(n, n))[1], col)  # x[n*cab + cust] layout (procedure.py:56)
    return st


def test_assign_kats(td):
    gold = load_golden("assign.json")
    for key in ("A1_python_py", "A2_glpk_mod", "A3_procedure_py"):
        _check_assign(td, np.array(gold[key]["cost"]), gold[key]["objective"])
    x = td.solve(4, gold["A1_python_py"]["cost"])
    perm = np.nonzero(np.asarray(x).reshape(4, 4))[1].tolist()
    assert perm in gold["A1_python_py"]["optimal_perms"]
    assert td.solve(0, []) == (0, [])                            # solver.py:12


def test_assign_config1(td):
    gold = load_golden("assign.json")
    _check_assign(td, g.config1a(), gold["config1a"]["objective"])
    # LP-relaxation clause of north_star: |objective - LP optimum of the reference model| <= 1e-9 relative
    assert abs(gold["config1a"]["lp"] - gold["config1a"]["objective"]) <= 1e-9 * gold["config1a"]["objective"]
    dist, cab_to, cust_from = g.config1b()
    n, cost = cost_ref.calculate_cost_np(dist, cab_to, cust_from)
    _check_assign(td, cost, gold["config1b"]["objective"])
    assert abs(gold["config1b"]["lp"] - gold["config1b"]["objective"]) <= 1e-9 * gold["config1b"]["objective"]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 64, 130, 257, 301])
def test_assign_random_small(td, n):
    rng = np.random.default_rng(100 + n)
    _check_assign(td, rng.integers(1, 40, (n, n)))
    _check_assign(td, rng.integers(0, 3, (n, n)))               # brutal ties
    _check_assign(td, rng.integers(0, 250001, (n, n)))          # wide range
    _check_assign(td, np.full((n, n), 7))                        # all equal
    _check_assign(td, rng.integers(-50, 50, (n, n)))            # negative costs


@pytest.mark.parametrize("n", [4, 8, 128, 132, 256, 260, 508, 1024])
def test_assign_ring_path_sizes(td, n):
    """n % 4 == 0 takes the cp.async ring sweeps: partial tiles, half-tiles, one and several tiles"""
    rng = np.random.default_rng(300 + n)
    _check_assign(td, rng.integers(1, 40, (n, n)))
    _check_assign(td, rng.integers(0, 3, (n, n)))
    _check_assign(td, rng.integers(-1000, 250001, (n, n)))
    _check_assign(td, np.abs(rng.integers(0, 50, n)[:, None] - rng.integers(0, 50, n)[None, :]))   # stand-derived, degenerate


@pytest.mark.parametrize("n", [64, 200, 301])
def test_assign_large_magnitudes_take_the_64_bit_sweeps(td, n):
    """4 max|c| + 2 sum(D) + level >= 2^30 switches a level to the 64-bit relaxation; max|c| >= 2^29 also the init"""
    rng = np.random.default_rng(400 + n)
    _check_assign(td, rng.integers(0, 2**28 + 12345, (n, n)))          # 32-bit init, 64-bit relaxation
    _check_assign(td, rng.integers(-2**30, 2**30, (n, n)))             # 64-bit everywhere
    c = rng.integers(1, 40, (n, n))
    c[rng.integers(0, n, 5)] = 2**30                                    # a few huge dummy rows
    _check_assign(td, c)


def test_assign_small_instances_both_kernels(td, monkeypatch):
    """n <= 232 balanced instances run in one CTA (assign_small_kernel, everything in shared memory); the cooperative
    kernel must keep solving them too (TD_ASSIGN_NO_SMALL=1).  Objective vs scipy, permutation, and the dual certificate
    of either kernel, on degenerate (U[0,2]), uniform, negative and 2^30-magnitude costs and on the KATs."""
    rng = np.random.default_rng(88)
    cases = [np.array([5, 5, 0, 5, 1, 1, 3, 8, 9, 9, 5, 0, 100, 100, 100, 100]).reshape(4, 4), g.config1a()]
    for n in (1, 2, 3, 7, 33, 100, 199, 232):
        cases.append(rng.integers(0, 3, (n, n)))
        cases.append(rng.integers(-1000, 1000, (n, n)))
    cases.append(rng.integers(0, 1 << 30, (150, 150)))
    cases.append(np.full((50, 50), 7))
    for no_small in ("", "1"):
        if no_small:
            monkeypatch.setenv("TD_ASSIGN_NO_SMALL", "1")
        for M in cases:
            M = np.ascontiguousarray(M, dtype=np.int32)
            _check_assign(td, M)
            obj, col, u, v, cert = _certify(td, M)
            assert cert["optimal"] and obj == cert["dual_objective"], (no_small, M.shape, cert)
    monkeypatch.delenv("TD_ASSIGN_NO_SMALL", raising=False)


def test_assign_is_deterministic(td):
    dist, cab_to, cust_from = g.config1b()
    n, cost = cost_ref.calculate_cost_np(dist, cab_to, cust_from)
    cols = [td.solve_full(n, cost)[1].tolist() for _ in range(3)]
    assert cols[0] == cols[1] == cols[2]
    c2 = g.config2_stand()
    a = td.solve_full(c2.shape[0], c2)[1]
    b = td.solve_full(c2.shape[0], c2)[1]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n_cabs,n_cust,cutoff", [(600, 218, 10), (218, 600, 10), (351, 600, None), (600, 599, None),
                                                  (1, 40, None), (40, 1, None), (130, 257, 10), (1300, 700, 10),
                                                  (0, 12, None), (12, 0, None)])
def test_assign_unbalanced_native(td, n_cabs, n_cust, cutoff):
    """only the real block of a padded instance is searched; objective and layout are unchanged"""
    import torch
    rng = np.random.default_rng(n_cabs * 1000 + n_cust)
    dist = g.stand_distances(50)
    n, cost = cost_ref.calculate_cost_np(dist, rng.integers(0, 50, n_cabs), rng.integers(0, 50, n_cust), cutoff=cutoff)
    ref_obj = assign_ref.solve_scipy(cost)[0]
    eng = td.engine()
    c = torch.from_numpy(np.ascontiguousarray(cost, dtype=np.int32)).cuda()
    col, obj, x, st = eng.assign(c, want_x=True, want_stats=True, n_real_rows=n_cabs if n_cabs < n else None,
                                 n_real_cols=n_cust if n_cust < n else None)
    col, x = col.cpu().numpy(), x.cpu().numpy()
    assert int(obj.item()) == ref_obj
    assert sorted(col.tolist()) == list(range(n)) and assign_ref.check_x(x, n)
    assert int(cost[np.arange(n), col].sum()) == ref_obj
    assert np.array_equal(np.nonzero(x.reshape(n, n))[1], col)
    # the balanced entry point on the same matrix agrees on the objective
    assert int(eng.assign(c)[1].item()) == ref_obj
    if n <= 300:   # host-buffer twin (what a cgo / JNI caller binds)
        import ctypes
        from taxidispatcher_b200 import _lib
        hc = np.ascontiguousarray(cost, dtype=np.int32)
        hcol, hobj = np.empty(n, np.int32), ctypes.c_int64()
        rc = _lib.lib().tdh_assign_exact_rect(hc.ctypes.data_as(ctypes.c_void_p), n, min(n_cabs, n), min(n_cust, n),
                                              hcol.ctypes.data_as(ctypes.c_void_p), ctypes.byref(hobj), None, None)
        assert rc == 0 and hobj.value == ref_obj and hcol.tolist() == col.tolist()
    # reference-shaped wrapper (split.py:139 signature) goes through the same path
    cabs = [(i, 0, int(t)) for i, t in enumerate(rng.integers(0, 50, min(n_cabs, 64)))]
    dem = [(i, int(f), 0) for i, f in enumerate(rng.integers(0, 50, min(n_cust, 48)))]
    nn, xx, cc = td.solve_dispatch(dist, dem, cabs)
    assert int((np.asarray(cc).reshape(-1) * np.asarray(xx).reshape(-1)).sum()) == assign_ref.solve_scipy(np.asarray(cc))[0]


def test_assign_structured(td):
    rng = np.random.default_rng(77)
    for n_cabs, n_cust, S, cutoff in ((120, 200, 50, None), (200, 120, 50, 10), (218, 600, 50, 10), (600, 351, 50, 10)):
        dist = g.stand_distances(S)
        n, cost = cost_ref.calculate_cost_np(dist, rng.integers(0, S, n_cabs), rng.integers(0, S, n_cust), cutoff=cutoff)
        _check_assign(td, cost)
    _check_assign(td, np.abs(np.arange(150)[:, None] - np.arange(150)[None, :]))   # identity optimum, objective 0
    _check_assign(td, (np.arange(90)[:, None] * np.arange(90)[None, :]) % 17)


def test_assign_config2_and_mid_sizes(td):
    gold = load_golden("assign.json")
    _check_assign(td, g.config2(), gold["config2"]["objective"])
    _check_assign(td, g.config2_stand(), gold["config2_stand"]["objective"])
    _check_assign(td, g.config5a(5000), gold["config5a_n5000"]["objective"])
    _check_assign(td, g.config5b_cost(5000), gold["config5b_n5000"]["objective"])


def test_assign_20k_north_star_objectives(td):
    """north star: exact optimum of 20 000 x 20 000.  For costs |a_i - b_j| the sorted matching is optimal (exchange
    argument), so the optimum of config 5-B is sum |sort(cab_to) - sort(cust_from)|; on config 5-A (U[1,39]) every row
    holds a 1 and the optimum reaches the lower bound n * 1 (checked as equality with the bound)."""
    import torch
    cab_to, cust_from = g.config5b()
    closed = int(np.abs(np.sort(cab_to).astype(np.int64) - np.sort(cust_from).astype(np.int64)).sum())
    assert closed == 480177
    eng = td.engine()
    a = torch.from_numpy(cab_to).cuda().to(torch.int32)
    b = torch.from_numpy(cust_from).cuda().to(torch.int32)
    cost = (a[:, None] - b[None, :]).abs().contiguous()
    col, obj, _, st = eng.assign(cost, want_stats=True)
    col_h = col.cpu().numpy().astype(np.int64)
    assert sorted(col_h.tolist()) == list(range(20000))
    assert int(obj.item()) == closed == int(np.abs(cab_to.astype(np.int64) - cust_from[col_h]).sum())
    del cost
    c5a = torch.from_numpy(g.config5a()).cuda()
    col, obj, _, _ = eng.assign(c5a)
    col_h = col.cpu().numpy().astype(np.int64)
    assert sorted(col_h.tolist()) == list(range(20000))
    assert int(obj.item()) == 20000 == int(c5a.cpu().numpy()[np.arange(20000), col_h].sum())


def _certify(td, cost, nr=None, nc=None):
    """solve + dual certificate; returns (objective, certificate dict)"""
    import torch
    eng = td.engine()
    c = torch.from_numpy(np.ascontiguousarray(cost, dtype=np.int32)).cuda() if isinstance(cost, np.ndarray) else cost
    n = int(c.shape[0])
    col, obj, _, _ = eng.assign(c, n_real_rows=nr, n_real_cols=nc)
    u, v = eng.assign_duals(n, nr, nc)
    cert = eng.assign_certify(c, col, u, v, nr, nc)
    return int(obj.item()), col, u, v, cert


def test_assign_dual_certificate(td):
    """The potentials the solver ends with certify its matching (complementary slackness): feasible on the real block,
    tight on the matched cells, dual objective == primal.  Checked against scipy where scipy is fast, and the checker
    itself is shown to reject a perturbed matching / perturbed potentials."""
    import torch
    rng = np.random.default_rng(61)
    for M in (g.config1a(), rng.integers(-50, 50, (301, 301)), rng.integers(0, 3, (260, 260)), g.config2_stand()[:1024, :1024]):
        M = np.ascontiguousarray(M, dtype=np.int32)
        obj, col, u, v, cert = _certify(td, M)
        assert cert["optimal"], cert
        assert obj == cert["dual_objective"] == assign_ref.solve_scipy(M)[0]
    # unbalanced instances: padding rows (more customers than cabs) and padding columns (more cabs than customers)
    for n_cabs, n_cust in ((218, 600), (600, 218), (351, 600), (64, 65)):
        S = 50
        dist = g.stand_distances(S)
        cab_to = rng.integers(0, S, n_cabs)
        cust_from = rng.integers(0, S, n_cust)
        n, cost = cost_ref.calculate_cost_np(dist, cab_to.astype(np.int32), cust_from.astype(np.int32))
        nr = n_cabs if n_cabs < n else None
        nc = n_cust if n_cust < n else None
        obj, col, u, v, cert = _certify(td, cost, nr, nc)
        assert cert["optimal"], (n_cabs, n_cust, cert)
        pad = g.BIG_COST * (n - min(n_cabs, n_cust))
        assert obj == cert["matched_real_cost"] + pad == assign_ref.solve_scipy(cost)[0]
    # the checker rejects what is not optimal
    M = np.ascontiguousarray(g.config1a(), dtype=np.int32)
    obj, col, u, v, cert = _certify(td, M)
    eng = td.engine()
    Md = torch.from_numpy(M).cuda()
    col2 = col.clone()
    col2[[0, 1]] = col[[1, 0]]                                       # another permutation: some matched cell is not tight
    bad = eng.assign_certify(Md, col2, u, v)
    assert not bad["optimal"] and (bad["max_matched_slack"] > 0 or bad["dual_objective"] != bad["matched_real_cost"])
    u2 = u.clone()
    u2[5] += 1                                                       # infeasible potentials
    assert not eng.assign_certify(Md, col, u2, v)["optimal"]


def test_assign_20k_certificates(td):
    """north star: the exact optimum of 20 000 x 20 000 instances is PROVED by the dual certificate (one sweep) --
    config 5-B (closed form known as well), config 5-A, and a random instance with no closed form at all."""
    import torch
    cab_to, cust_from = g.config5b()
    a = torch.from_numpy(cab_to).cuda().to(torch.int32)
    b = torch.from_numpy(cust_from).cuda().to(torch.int32)
    cost = (a[:, None] - b[None, :]).abs().contiguous()
    obj, col, u, v, cert = _certify(td, cost)
    assert cert["optimal"] and obj == cert["dual_objective"] == 480177, cert
    del cost
    obj, col, u, v, cert = _certify(td, torch.from_numpy(g.config5a()).cuda())
    assert cert["optimal"] and obj == cert["dual_objective"] == 20000, cert
    gen = torch.Generator(device="cuda").manual_seed(20002)
    cost = torch.randint(0, 100000, (20000, 20000), generator=gen, device="cuda", dtype=torch.int32)
    obj, col, u, v, cert = _certify(td, cost)
    assert cert["optimal"] and obj == cert["dual_objective"] == cert["matched_real_cost"], cert
    assert sorted(col.cpu().numpy().tolist()) == list(range(20000))


def test_assign_optimum_not_above_lcm(td):
    """heuristic.py:40's invariant: the optimum never exceeds the LCM total"""
    C = g.config2()[:400, :400].copy()
    _, _, obj, _ = td.solve_full(400, C)
    assert obj <= td.LCM_heuristic(400, C)


def test_pool_batched_shards_match_single_shard_calls(td):
    """td_pool_find_shards: several consecutive logical shards in one enumeration + one selection launch"""
    gold = load_golden("pool722.json")
    dem = g.pool_demand()
    dist = g.stand_distances(50)
    for begin, count in ((0, 8), (2, 3), (7, 1)):
        res = td.find_pool_block(dem, dist, 4, begin, count, 8)
        assert len(res) == count
        for s, (plans, st) in enumerate(res):
            gs = gold["shards"][begin + s]
            assert plans.tolist() == gs["plans"], (begin, s)
            assert {q: st[q] for q in ("evaluated", "feasible", "kept")} == gs["stats"]
    merged, stats = td.find_pool_all(dem, dist, 4)
    assert merged.tolist() == gold["merged"]
    assert stats["evaluated"] == sum(g.POOL722_EVALUATED) and stats["kept_per_shard"] == g.POOL722_KEPT
    # ragged: more shards than leaders, pool size 3, odd shard count
    dem2 = g.pool_demand(45, seed=4)
    d51 = g.stand_distances(51)
    for n_shards in (5, 64):
        m, st = td.find_pool_all(dem2, d51, 3, n_shards=n_shards)
        ref = pool_ref.merge([pool_ref.find(dem2, d51, 3, sh, n_shards)[0] for sh in range(n_shards)], 45, 3)
        assert np.array_equal(m, ref), n_shards


# ---- file / CLI protocols ------------------------------------------------------------
def test_cli_twins_speak_the_reference_file_protocols(td, tmp_path, monkeypatch):
    import os
    from taxidispatcher_b200 import formats
    from taxidispatcher_b200.cli import findpool, pool_n, solver
    case = [c for c in load_golden("pool_small.json") if c["label"] == "n180_seed11_k4"][0]
    dem = np.array(case["demand"], dtype=np.int32)
    monkeypatch.chdir(tmp_path)
    (tmp_path / "demand.csv").write_text(formats.write_demand_csv(dem))
    # pool_n <pool-size> <thread> <demand-file> <rec-number> <output-file>   (pool_n.c:211-219)
    for sh in (0, 3, 7):
        assert pool_n.main(["4", str(sh), "demand.csv", str(len(dem)), "out%d.csv" % sh]) == 0
        assert os.path.exists("out%d.flg" % sh)                                   # pool_n.c:56-62
        got = formats.read_result_csv((tmp_path / ("out%d.csv" % sh)).read_text(), 4)
        assert got.tolist() == case["shards"][sh]["plans"]
        assert (tmp_path / ("out%d.csv" % sh)).read_text() == pool_ref.format_result_csv(np.array(case["shards"][sh]["plans"]), 4)
    assert pool_n.main(["4", "0", "missing.csv", "10", "o.csv"]) == 1            # pool_n.c:36-39
    assert pool_n.main(["4", "0"]) == 1                                           # usage
    # findpool <pool-size> <demand-file> <rec-number> <output-file>           (findpool.c:127-134)
    assert findpool.main(["4", "demand.csv", str(len(dem)), "pool_out.csv"]) == 0
    merged = formats.read_result_csv((tmp_path / "pool_out.csv").read_text(), 4)
    assert merged[:, :8].tolist() == [r[:8] for r in case["merged"]]
    # solver.py: cost.txt -> solv_out.txt                                      (solver.py:30-39)
    C = g.config1a(60)
    (tmp_path / "cost.txt").write_text(formats.write_cost_txt(C))
    assert solver.main(["cost.txt", "solv_out.txt"]) == 0
    x = formats.read_solv_out((tmp_path / "solv_out.txt").read_text(), 60)
    assert assign_ref.check_x(x, 60)
    assert int((C.reshape(-1) * x).sum()) == assign_ref.solve_scipy(C)[0]


def test_pool_multi_pass_cost_windows(td):
    """A record capacity far below the feasible count forces the cost-window passes (later windows
    enumerate only customers that are still free); results must not change."""
    import torch
    gold = load_golden("pool722.json")
    eng = td.engine()
    dem = torch.from_numpy(g.pool_demand()).cuda()
    dist = torch.from_numpy(g.stand_distances(50)).cuda()
    out, cnt, st = eng.pool_find_shards(dem, dist, 4, 0, 8, 8, max_feasible=600_000)   # ~16 M feasible plans in total
    counts = cnt.cpu().numpy()
    assert max(s.passes for s in st) >= 3
    for s in range(8):
        assert out[s, : counts[s]].cpu().numpy().tolist() == gold["shards"][s]["plans"], s
        assert (st[s].evaluated, st[s].feasible, st[s].kept) == tuple(gold["shards"][s]["stats"][q] for q in ("evaluated", "feasible", "kept"))
    # small ragged case, pool sizes 2..4, tiny capacity
    case_dem = np.array(load_golden("pool_small.json")[3]["demand"], dtype=np.int32)
    d51 = g.stand_distances(51)
    for k in (2, 3, 4):
        for sh in (0, 5):
            o, c, s2 = eng.pool_find_shards(torch.from_numpy(case_dem).cuda(), torch.from_numpy(d51).cuda(), k, sh, 1, 8, max_feasible=64)
            ref, rst = pool_ref.find(case_dem, d51, k, sh)
            assert o[0, : int(c[0])].cpu().numpy().tolist() == ref.tolist(), (k, sh)
            assert (s2[0].evaluated, s2[0].feasible) == (rst["evaluated"], rst["feasible"])


def test_pool_1200_customers_golden_single_pass_and_windows(td):
    """1200 customers (2.5e10 leaf plans, 1.5e8 feasible): the compiled reference needed minutes per shard; its
    32-bit count_all wraps (pool_n.c:28), so evaluated is compared modulo 2^32.  Run once with room for every
    record and once with a record list 40x too small (cost windows + alive pruning)."""
    import torch
    gold = load_golden("pool1200.json")
    dem = g.pool_demand(1200, seed=1200)
    assert sha(dem) == gold["demand_sha256"]
    eng = td.engine()
    dd, ds = torch.from_numpy(dem).cuda(), torch.from_numpy(g.stand_distances(50)).cuda()
    for mf in (200_000_000, 4_000_000):
        out, cnt, st = eng.pool_find_shards(dd, ds, 4, 0, 8, 8, max_feasible=mf)
        counts = cnt.cpu().numpy()
        for s in range(8):
            gs = gold["shards"][s]
            assert out[s, : counts[s]].cpu().numpy().tolist() == gs["plans"], (mf, s)
            assert st[s].feasible == gs["stats"]["feasible"] and st[s].kept == gs["stats"]["kept"]
            assert (st[s].evaluated - gs["stats"]["evaluated"]) % (1 << 32) == 0
        merged, mc = eng.pool_merge_padded(out, cnt, None, 1200, 4)
        assert merged[: int(mc.item())].cpu().numpy().tolist() == gold["merged"]
    assert st[0].passes > 1


def test_pool_merge_padded_sorted_unsorted_and_many_slots(td):
    """td_pool_merge_padded ranks by binary search when every slot is an ascending run (what pool_emit produces), sorts
    when a caller hands in rows in another order, and takes the arrival-order compaction above 64 slots: all three must
    equal the restated merge (findpool.c:83-108) of the same rows."""
    import torch
    eng = td.engine()
    dem = g.pool_demand(300, seed=77)
    dist = g.stand_distances(50)
    dd, ds = torch.from_numpy(dem).cuda(), torch.from_numpy(dist).cuda()
    rng = np.random.default_rng(5)
    for n_shards in (8, 70):
        outs, cnts = [], []
        for b in range(0, n_shards, 64):
            c = min(64, n_shards - b)
            out, cnt, st = eng.pool_find_shards(dd, ds, 4, b, c, n_shards)
            outs.append(out.clone()); cnts.append(cnt.clone())
        out, cnt = torch.cat(outs, 0), torch.cat(cnts, 0)
        counts = cnt.cpu().numpy()
        host = out.cpu().numpy()
        want = pool_ref.merge([host[s_, : counts[s_]] for s_ in range(n_shards)], 300, 4)
        merged, mc = eng.pool_merge_padded(out, cnt, None, 300, 4)
        assert merged[: int(mc.item())].cpu().numpy().tolist() == want.tolist(), n_shards
        # the same rows with one slot's rows permuted: the position in the input breaks ties, so the answer is the merge of
        # the permuted rows
        big = int(np.argmax(counts))
        perm = rng.permutation(int(counts[big]))
        host2 = host.copy()
        host2[big, : counts[big]] = host[big, : counts[big]][perm]
        want2 = pool_ref.merge([host2[s_, : counts[s_]] for s_ in range(n_shards)], 300, 4)
        merged2, mc2 = eng.pool_merge_padded(torch.from_numpy(host2).cuda(), cnt, None, 300, 4)
        assert merged2[: int(mc2.item())].cpu().numpy().tolist() == want2.tolist(), n_shards


def test_pool_headed_blocks_and_steady_state_path(td):
    """Headed blocks (survivors + count + counters in one buffer, the layout of the multi-GPU gather) and the merge on
    them: identical to the plain entry points, on config 3 and on a K = 2 / K = 3 case; find_pool_all takes the headed
    steady-state path from its second call on and must keep returning the golden result and the reference's counters."""
    import torch
    gold = load_golden("pool722.json")
    eng = td.engine()
    dem_np, dist_np = g.pool_demand(), g.stand_distances(50)
    from taxidispatcher_b200 import dispatch
    for rep in range(6):       # 1: cost-window path, 2-3: asynchronous headed path, 4-6: the captured graph (PoolJobGraph)
        merged, st = td.find_pool_all(dem_np, dist_np, 4)
        assert merged.tolist() == gold["merged"], rep
        assert st["evaluated"] == sum(g.POOL722_EVALUATED) and st["feasible"] == sum(g.POOL722_FEASIBLE)
        assert st["kept_per_shard"] == g.POOL722_KEPT and st["kept"] == 110
    gkey = ("all", 722, 50, 4, 8, eng.device.index)
    assert isinstance(dispatch._POOL_GRAPHS.get(gkey), dispatch.PoolJobGraph)
    # another input of the same shape through the same graph, checked against the piecewise path
    dem2 = g.pool_demand(722, seed=9)
    got, st2 = td.find_pool_all(dem2, dist_np, 4)
    graph = dispatch._POOL_GRAPHS.pop(gkey, None)          # (None: dem2 outgrew the record list and dropped the graph)
    want, st_want = td.find_pool_all(dem2, dist_np, 4)
    assert got.tolist() == want.tolist()
    assert {k_: st2[k_] for k_ in ("evaluated", "feasible", "kept_per_shard", "kept")} == \
           {k_: st_want[k_] for k_ in ("evaluated", "feasible", "kept_per_shard", "kept")}
    if graph is None:
        for rep in range(4):
            td.find_pool_all(dem_np, dist_np, 4)
        graph = dispatch._POOL_GRAPHS[gkey]
    dispatch._POOL_GRAPHS[gkey] = graph
    # a bigger job reallocates the engine's workspaces: the graph notices and the next call rebuilds its way up again
    td.find_pool_all(g.pool_demand(900, seed=3), dist_np, 4)
    ok_before = graph.valid()
    merged, st = td.find_pool_all(dem_np, dist_np, 4)
    assert merged.tolist() == gold["merged"] and (ok_before or dispatch._POOL_GRAPHS.get(gkey) is not graph)
    dem, dist = torch.from_numpy(dem_np).cuda(), torch.from_numpy(dist_np).cuda()
    cap = 722 // 2 + 1
    # two "ranks" with 5 + 3 shards, padded to 5 slots each, gathered in rank order: the multi-GPU layout on one device
    slots = 5
    all_blocks = torch.zeros((2 * slots, cap + 1, 9), dtype=torch.int32, device="cuda")
    eng.pool_find_shards(dem, dist, 4, 0, 5, 8)                                  # sizes the record list for 5 shards
    eng.pool_find_shards_headed(dem, dist, 4, 0, 5, 8, out=all_blocks[0:5])
    eng.pool_find_shards(dem, dist, 4, 5, 3, 8)
    eng.pool_find_shards_headed(dem, dist, 4, 5, 3, 8, out=all_blocks[5:8])
    slot_shard = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7, 0, 0], dtype=torch.int32, device="cuda")
    plans, counts, ev, fe = eng.pool_merge_headed_packed(all_blocks, slot_shard, 722, 4)
    assert plans.tolist() == gold["merged"]
    assert counts.tolist() == g.POOL722_KEPT + [0, 0]
    assert ev.tolist()[:8] == g.POOL722_EVALUATED and fe.tolist()[:8] == g.POOL722_FEASIBLE
    for s_ in range(8):
        assert all_blocks[s_, 1: 1 + int(counts[s_])].cpu().numpy().tolist() == gold["shards"][s_]["plans"]
    # small ragged case, every pool size (K < 4 keeps the reference's concatenation-order quirk in the merge)
    case = load_golden("pool_small.json")
    for c in case[6:12]:
        d_np = np.array(c["demand"], dtype=np.int32)
        for rep in range(5):                       # the last two calls run the captured graph
            merged, st = td.find_pool_all(d_np, g.stand_distances(c["n_stands"]), c["pool_size"])
            assert merged.tolist() == c["merged"], (c["label"], rep)


def test_pool_async_overflow_reports_minus_one_and_recovers(td):
    """The asynchronous single pass (stats == NULL) cannot fall back to cost windows: when the record list is too small
    every count comes back as -1, nothing is read or written out of bounds (the selection does not run), and the
    same workspace serves a correct synchronous call afterwards."""
    import torch
    gold = load_golden("pool722.json")
    eng = td.engine()
    dem = torch.from_numpy(g.pool_demand()).cuda()
    dist = torch.from_numpy(g.stand_distances(50)).cuda()
    small = g.pool_demand(97, seed=13)
    eng.pool_find_shards(torch.from_numpy(small).cuda(), torch.from_numpy(g.stand_distances(51)).cuda(), 3, 0, 3, 3)  # other shape first
    out, cnt, token = eng.pool_find_shards(dem, dist, 4, 0, 8, 8, max_feasible=50_000, defer_stats=True)
    st, overflowed = eng.pool_read_stats(token)
    assert overflowed and (cnt.cpu().numpy() == -1).all()
    assert [int(x.evaluated) for x in st] == g.POOL722_EVALUATED          # the counters stay exact
    out, cnt, st = eng.pool_find_shards(dem, dist, 4, 0, 8, 8, max_feasible=50_000)
    counts = cnt.cpu().numpy()
    for s_ in range(8):
        assert out[s_, : counts[s_]].cpu().numpy().tolist() == gold["shards"][s_]["plans"], s_


def test_pool_5000_customers_reference_slices(td):
    """north-star size: 5000 waiting customers.  Three slices (10 leading customers each) of the reference's own shard
    rule with 512 shards were run through the compiled pool_n.c (tests/golden/make_golden.py --large-slices, ~10
    CPU-minutes per slice); plans, feasible and kept counts must be identical, evaluated modulo 2^32 (pool_n.c:28).
    Checked on the single-pass path and with a record list 3x too small (cost windows + alive pruning; a list that
    cannot hold one cost level is grown by the host wrapper, so it must not be too small either)."""
    import torch
    gold = load_golden("pool5000_slices.json")
    dem = g.pool_demand(gold["n"], seed=gold["seed"])
    assert sha(dem) == gold["demand_sha256"]
    eng = td.engine()
    dd, ds = torch.from_numpy(dem).cuda(), torch.from_numpy(g.stand_distances(gold["n_stands"])).cuda()
    for mf in (120_000_000, 20_000_000):
        for sl in gold["slices"]:
            out, cnt, st = eng.pool_find_shards(dd, ds, 4, sl["slice"], 1, gold["n_shards"], max_feasible=mf)
            m = int(cnt[0])
            assert out[0, :m].cpu().numpy().tolist() == sl["plans"], (mf, sl["slice"])
            assert st[0].feasible == sl["stats"]["feasible"] and st[0].kept == sl["stats"]["kept"]
            assert st[0].evaluated % (1 << 32) == sl["stats"]["evaluated_mod_2_32"]
            assert (st[0].passes > 1) == (mf < sl["stats"]["feasible"])


def test_pool_large_tables_take_the_global_memory_paths(td):
    """> 128 stands (distance table not staged in shared memory) and > 6144 customers (customer records not
    staged): the enumeration falls back to read-only global loads; results must not change."""
    rng = np.random.default_rng(77)
    S = 200
    dist = g.stand_distances(S)
    n = 260
    fr = rng.integers(0, S, n)
    to = (fr + rng.integers(1, 30, n)) % S
    dem = np.stack([np.arange(n), fr, to, rng.integers(0, 25, n), rng.integers(0, 40, n)], axis=1).astype(np.int32)
    for k in (2, 3, 4):
        plans, st = td.find_pool(dem, dist, k, 1, 4)
        oplans, ost = pool_ref.find(dem, dist, k, 1, 4)
        assert {q: st[q] for q in ost} == ost and np.array_equal(plans, oplans), k
    n = 6400                                              # 16-byte customer records exceed the 96 KB staging budget
    fr = rng.integers(0, 50, n)
    to = (fr + rng.integers(1, 20, n)) % 50
    dem = np.stack([np.arange(n), fr, to, np.zeros(n, int), rng.integers(0, 3, n)], axis=1).astype(np.int32)
    d50 = g.stand_distances(50)
    plans, st = td.find_pool(dem, d50, 2, 5, 8)           # wait 0: partners share the pickup stand
    oplans, ost = pool_ref.find(dem, d50, 2, 5, 8)
    assert {q: st[q] for q in ost} == ost and np.array_equal(plans, oplans)


def test_cost_matrix_large_stand_table_without_staging(td):
    """a stand row that does not fit the double-buffered shared-memory stage (S > 12288) uses the direct path"""
    import torch
    rng = np.random.default_rng(3)
    S = 13000
    eng = td.engine()
    dist = torch.randint(0, 1000, (S, S), dtype=torch.int32, device="cuda")
    cab_to = torch.from_numpy(rng.integers(0, S, 37).astype(np.int32)).cuda()
    cust_from = torch.from_numpy(rng.integers(0, S, 53).astype(np.int32)).cuda()
    out = eng.cost_matrix(dist, cab_to, cust_from, fill=-7, cutoff=900).cpu().numpy()
    d = dist.cpu().numpy()
    ref = np.full((53, 53), -7, np.int32)
    blk = d[cab_to.cpu().numpy()[:, None], cust_from.cpu().numpy()[None, :]]
    ref[:37, :53] = np.where(blk < 900, blk, -7)
    assert np.array_equal(out, ref)
