"""Variant: do not stop the Dijkstra at the first sink level; augment trees as their first sink appears,
keep going (dead trees keep growing) until a stop rule, then ONE dual update.  Count work."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch')
from oracle import gen_inputs as g, assign_ref, cost_ref
import proto_assign as P
INF = P.INF

def ssp_phases2(C, u, v, mate_r, mate_c, stop_frac=1.0, max_extra_levels=10**9, scan_dead=True):
    n = C.shape[0]
    st = dict(phases=0, levels=0, rows_scanned=0, augment=0)
    while True:
        free_r = np.nonzero(mate_r < 0)[0]
        if len(free_r) == 0: break
        st['phases'] += 1
        nfree0 = len(free_r)
        dist = np.full(n, INF, np.int64); pred = -np.ones(n, np.int64)
        settled = np.zeros(n, bool)
        d_row = np.full(n, INF, np.int64); root = -np.ones(n, np.int64)
        d_row[free_r] = 0; root[free_r] = free_r
        dead = np.zeros(n, bool)   # indexed by root row
        frontier = free_r
        naug = 0; first_sink_level = None; lvl = 0
        delta = 0
        while True:
            if len(frontier):
                st['rows_scanned'] += len(frontier)
                R = C[frontier].astype(np.int64) - u[frontier, None] - v[None, :] + d_row[frontier, None]
                R[:, settled] = INF
                best = R.min(0); arg = R.argmin(0)
                upd = best < dist
                dist[upd] = best[upd]; pred[upd] = frontier[arg[upd]]
            cand = np.where(settled, INF, dist)
            dmin = cand.min()
            if dmin >= INF: break            # every column settled
            delta = dmin
            st['levels'] += 1; lvl += 1
            newc = np.nonzero(cand == delta)[0]
            settled[newc] = True
            is_free = mate_c[newc] < 0
            sinks = newc[is_free]
            matched = newc[~is_free]
            frontier = mate_c[matched]
            d_row[frontier] = delta
            root[frontier] = root[pred[matched]]
            if not scan_dead:
                pass
            # augment alive trees (one sink per root, smallest col)
            for j in sinks:
                rt = root[pred[j]]
                if dead[rt]: continue
                dead[rt] = True
                cj = j
                while True:
                    i = pred[cj]; nxt = mate_r[i]
                    mate_r[i] = cj; mate_c[cj] = i
                    if nxt < 0: break
                    cj = nxt
                naug += 1; st['augment'] += 1
            if len(sinks) and first_sink_level is None: first_sink_level = lvl
            if naug >= max(1, int(stop_frac * nfree0)): break
            if first_sink_level is not None and lvl - first_sink_level >= max_extra_levels: break
        reached = d_row < INF
        u[reached] += delta - d_row[reached]
        v[settled] -= delta - dist[settled]
    return st

def run(name, C, **kw):
    n = C.shape[0]
    t = time.time()
    u, v = P.init_reduce(C)
    mr, mc = P.greedy_tight(C, u, v)
    f0 = int((mr < 0).sum())
    st = ssp_phases2(C, u, v, mr, mc, **kw)
    obj = int(C[np.arange(n), mr].sum())
    # verify duals
    red = C.astype(np.int64) - u[:, None] - v[None, :]
    okd = bool((red >= 0).all() and (red[np.arange(n), mr] == 0).all())
    ref = assign_ref.solve_scipy(C)[0]
    print(f"{name} {kw}: n={n} free0={f0} phases={st['phases']} levels={st['levels']} rows={st['rows_scanned']} ({st['rows_scanned']/n:.1f} sweeps) obj_ok={obj==ref} duals_ok={okd} t={time.time()-t:.1f}s", flush=True)

if __name__ == '__main__':
    for kw in (dict(stop_frac=1.0), dict(stop_frac=0.5), dict(stop_frac=0.25), dict(stop_frac=1.0, max_extra_levels=3), dict(stop_frac=1.0, max_extra_levels=10)):
        run('2stand', g.config2_stand(), **kw)
        run('5b-1000', g.config5b_cost(1000, 200), **kw)
    dist, cab_to, cust_from = g.config1b(); n, c = cost_ref.calculate_cost_np(dist, cab_to, cust_from); run('1b', c)
    run('1a', g.config1a())
