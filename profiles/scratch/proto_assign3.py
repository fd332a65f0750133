"""Jacobi augmenting-row-reduction (= eps-0 auction rounds) as warm start, then the forest SSP finish."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch')
from oracle import gen_inputs as g, assign_ref, cost_ref
import proto_assign as P
INF = P.INF

def arr_rounds(C, u, v, mate_r, mate_c, rounds, eps=0):
    n = C.shape[0]
    scanned = 0
    hist = []
    for r in range(rounds):
        free = np.nonzero(mate_r < 0)[0]
        hist.append(len(free))
        if len(free) == 0: break
        scanned += len(free)
        R = C[free].astype(np.int64) - v[None, :]
        j1 = R.argmin(1)
        u1 = R[np.arange(len(free)), j1]
        R2 = R.copy(); R2[np.arange(len(free)), j1] = INF
        u2 = R2.min(1) if n > 1 else u1
        inc = u2 - u1 + eps
        # winner per column: largest inc, ties lowest row
        order = np.lexsort((free, -inc, j1))   # sort by j1, then -inc, then row
        fj = j1[order]; first = np.ones(len(order), bool); first[1:] = fj[1:] != fj[:-1]
        win = order[first]
        wi = free[win]; wj = j1[win]; winc = inc[win]
        # only useful if column free or increment > 0 (otherwise pure swap with no progress)
        useful = (mate_c[wj] < 0) | (winc > 0)
        wi, wj, winc, w_u2 = wi[useful], wj[useful], winc[useful], u2[win][useful]
        old = mate_c[wj]
        mate_r[old[old >= 0]] = -1
        mate_c[wj] = wi; mate_r[wi] = wj
        v[wj] -= winc
        u[wi] = w_u2 + (eps if False else 0)
    return scanned, hist

def run(name, C, rounds, verify=True):
    n = C.shape[0]
    t = time.time()
    u, v = P.init_reduce(C)
    mr, mc = P.greedy_tight(C, u, v)
    f0 = int((mr < 0).sum())
    sc, hist = arr_rounds(C, u, v, mr, mc, rounds)
    # re-derive u for feasibility: u_i = min_j(c_ij - v_j) for all rows (v only decreased) -- matched rows keep tightness?
    red = C.astype(np.int64) - u[:, None] - v[None, :]
    feas = bool((red >= 0).all()); m = mr >= 0
    tight = bool((red[np.nonzero(m)[0], mr[m]] == 0).all())
    f1 = int((mr < 0).sum())
    st = P.ssp_phases(C, u, v, mr, mc, verbose=False)
    obj = int(C[np.arange(n), mr].sum())
    ref = assign_ref.solve_scipy(C)[0] if verify else obj
    print(f"{name} arr_rounds={rounds}: n={n} free0={f0} free_after_arr={f1} arr_scans={sc} ({sc/n:.1f} sw) feas={feas} tight={tight} phases={st['phases']} levels={st['levels']} rows={st['rows_scanned']} ({st['rows_scanned']/n:.1f} sw) ok={obj==ref} t={time.time()-t:.1f}s hist={hist[:12]}..{hist[-3:]}", flush=True)

if __name__ == '__main__':
    for rounds in (0, 10, 50, 200):
        run('2stand', g.config2_stand(), rounds)
        run('5b-1000', g.config5b_cost(1000, 200), rounds)
    for rounds in (50, 200):
        run('5b-5000', g.config5b_cost(5000, 1000), rounds)
        run('5a-5000', g.config5a(5000), rounds)
