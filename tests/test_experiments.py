"""split.py / greedy_opt.py orchestration.  CPU tests inject oracle solvers; GPU tests use the engine.
The optimum of every sub-instance is unique, which pins greedy_then_solve completely; solve_split additionally depends
on WHICH optimum the solver returns for the ranges (the unserved ids feed the last solve), so it is checked through
invariants and against the same restatement fed with the same solver."""
import numpy as np
import pytest

from oracle import assign_ref, cost_ref, gen_inputs as g, lcm_ref
from taxidispatcher_b200 import experiments as ex


def oracle_solver(distances, demand, cabs):
    n, cost = cost_ref.calculate_cost(distances, demand, cabs)
    if n == 0:
        return 0, [], 0
    _, col = assign_ref.solve_scipy(np.asarray(cost))
    return n, assign_ref.x_from_cols(col), np.asarray(cost)


def oracle_lcm(n, c, threshold):
    tot, sup, dem = lcm_ref.lcm_greedy_opt(n, c, threshold)
    return int(tot), sup, dem


def instance(seed, n_size=60, n_stands=20):
    rng = np.random.default_rng(seed)
    return g.stand_distances(n_stands), g.rand_list(rng, n_size, n_stands), g.rand_list(rng, n_size, n_stands)


def global_optimum(dist, demand, cabs):
    n, x, cost = oracle_solver(dist, demand, cabs)
    return ex.count_sum(n, cost, x, demand, cabs, dist, 0)


def test_split_invariants_cpu():
    for seed in range(6):
        dist, demand, cabs = instance(seed)
        total = ex.solve_split(20, dist, demand, cabs, solver=oracle_solver)
        opt = global_optimum(dist, demand, cabs)
        assert total >= opt                       # splitting can only lose (PDF p.5-6: 16 % on average)
    assert ex.solve_split(20, g.stand_distances(20), [], [(0, 1, 2)], solver=oracle_solver) is None   # split.py:62-64
    # every customer and cab in ONE range: the split solution is the global optimum
    dist = g.stand_distances(20)
    demand = [(i, 1 + i % 3, 9) for i in range(7)]
    cabs = [(i, 0, 2 + i % 2) for i in range(7)]
    assert ex.solve_split(20, dist, demand, cabs, solver=oracle_solver) == global_optimum(dist, demand, cabs)


def test_greedy_then_solve_cpu_matches_reference_flow():
    dist, demand, cabs = instance(11, n_size=120, n_stands=40)
    nn, opt, n2, hybrid = ex.greedy_then_solve(dist, demand, cabs, threshold=3, solver=oracle_solver, lcm=oracle_lcm)
    assert nn == max(len(demand), len(cabs)) and hybrid >= opt and n2 <= nn
    assert opt == global_optimum(dist, demand, cabs)


@pytest.mark.gpu
def test_experiments_on_gpu(td):
    for seed in (1, 2, 3):
        dist, demand, cabs = instance(seed, n_size=150, n_stands=40)
        # greedy_opt: LCM trace and both optima are unique -> bit-exact against the oracle flow
        got = ex.greedy_then_solve(dist, demand, cabs, threshold=3)
        ref = ex.greedy_then_solve(dist, demand, cabs, threshold=3, solver=oracle_solver, lcm=oracle_lcm)
        assert got == ref
        total = ex.solve_split(40, dist, demand, cabs)
        assert total >= global_optimum(dist, demand, cabs)
        # the same restatement fed with the GPU's own assignment vectors must agree with itself range by range
        assert total == ex.solve_split(40, dist, demand, cabs, solver=td.solve_dispatch)
