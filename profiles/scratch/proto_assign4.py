"""eps-scaling Jacobi auction (limited rounds) as warm start -> rounded feasible duals -> forest SSP finish."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch')
from oracle import gen_inputs as g, assign_ref, cost_ref
import proto_assign as P
INF = P.INF

def auction(C, K, eps_list, max_rounds_per_eps, stop_free_frac):
    """min-cost auction on costs C*K (integers). price p_j >= 0 (value = -(cK + p)). returns mate arrays, prices, scans"""
    n = C.shape[0]
    CK = C.astype(np.int64) * K
    p = np.zeros(n, np.int64)
    mate_r = -np.ones(n, np.int64); mate_c = -np.ones(n, np.int64)
    scans = 0; rounds = 0
    for eps in eps_list:
        # keep assignments that satisfy eps-CS? simplest: unassign everything at each new eps
        mate_r[:] = -1; mate_c[:] = -1
        for r in range(max_rounds_per_eps):
            free = np.nonzero(mate_r < 0)[0]
            if len(free) <= stop_free_frac * n and eps == eps_list[-1]: break
            if len(free) == 0: break
            rounds += 1; scans += len(free)
            R = CK[free] + p[None, :]
            j1 = R.argmin(1); a = np.arange(len(free))
            w1 = R[a, j1]
            R[a, j1] = INF
            w2 = R.min(1)
            bid = w2 - w1 + eps   # price increase
            order = np.lexsort((free, -bid, j1))
            fj = j1[order]; first = np.ones(len(order), bool); first[1:] = fj[1:] != fj[:-1]
            win = order[first]
            wi = free[win]; wj = j1[win]
            old = mate_c[wj]
            mate_r[old[old >= 0]] = -1
            mate_c[wj] = wi; mate_r[wi] = wj
            p[wj] += bid[win]
    return mate_r, mate_c, p, scans, rounds

def run(name, C, K, eps_list, max_rounds, stop_frac, verify=True):
    n = C.shape[0]
    t = time.time()
    mr, mc, p, scans, rounds = auction(C, K, eps_list, max_rounds, stop_frac)
    f1 = int((mr < 0).sum())
    # back to the original domain: v = -floor(p / K) (price up = v down); u = min_j (c - v)
    v = -(p // K)
    u = (C.astype(np.int64) - v[None, :]).min(1)
    m = np.nonzero(mr >= 0)[0]
    slackm = C[m, mr[m]].astype(np.int64) - u[m] - v[mr[m]]
    v[mr[m]] -= -(-slackm)  # v_j = c_ij - u_i  (lowering v keeps feasibility)
    v[mr[m]] = C[m, mr[m]].astype(np.int64) - u[m]
    red = C.astype(np.int64) - u[:, None] - v[None, :]
    feas = bool((red >= 0).all()); tight = bool((red[m, mr[m]] == 0).all())
    st = P.ssp_phases(C, u, v, mr, mc, verbose=False)
    obj = int(C[np.arange(n), mr].sum())
    ref = assign_ref.solve_scipy(C)[0] if verify else obj
    print(f"{name} K={K} eps={eps_list} maxr={max_rounds}: n={n} auction rounds={rounds} scans={scans} ({scans/n:.1f} sw) free_after={f1} forced_slack_sum={int(slackm.sum())} feas={feas} tight={tight} | phases={st['phases']} levels={st['levels']} rows={st['rows_scanned']} ({st['rows_scanned']/n:.1f} sw) ok={obj==ref} t={time.time()-t:.1f}s", flush=True)

if __name__ == '__main__':
    for name, C in (('2stand', g.config2_stand()), ('5b-1000', g.config5b_cost(1000, 200)), ('5b-5000', g.config5b_cost(5000, 1000))):
        n = C.shape[0]
        for K, eps_list, maxr in ((1, [1], 200), (4, [16, 4, 1], 200), (4, [64, 16, 4, 1], 500), (16, [256, 64, 16, 4, 1], 500)):
            run(name, C, K, eps_list, maxr, 0.01)
