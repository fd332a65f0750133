"""numpy prototype of the K2 exact finish (forest shortest-augmenting-path phases with integer
Dial levels) to count phases / levels / row scans before writing CUDA.  Not product code."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import gen_inputs as g, assign_ref, cost_ref

INF = np.iinfo(np.int64).max // 4
ROT = True
ROUNDS = 4

def init_reduce(C):
    n = C.shape[0]
    v = C.min(0).astype(np.int64)
    u = (C - v[None, :]).min(1).astype(np.int64)
    return u, v

def greedy_tight(C, u, v, rounds=None):
    rounds = rounds or ROUNDS
    """parallel-style greedy maximal matching on tight edges: each free row proposes to its first tight free col;
    each col accepts the lowest row; repeat."""
    n = C.shape[0]
    mate_r = -np.ones(n, np.int64); mate_c = -np.ones(n, np.int64)
    tight = (C - u[:, None] - v[None, :]) == 0
    for _ in range(rounds):
        free_r = np.nonzero(mate_r < 0)[0]
        if len(free_r) == 0: break
        T = tight[free_r] & (mate_c < 0)[None, :]
        has = T.any(1)
        if not has.any(): break
        if ROT:
            nfr = len(free_r)
            off = (free_r * 2654435761) % n
            # rotate each row by its offset: first tight free col at/after off (cyclic)
            idx = (np.arange(n)[None, :] + off[:, None]) % n
            Trot = np.take_along_axis(T, idx, axis=1)
            first = (Trot.argmax(1) + off) % n
        else:
            first = T.argmax(1)
        prop_rows = free_r[has]; prop_cols = first[has]
        # col accepts lowest row
        order = np.argsort(prop_rows, kind='stable')
        pr = prop_rows[order]; pc = prop_cols[order]
        uniq, idx = np.unique(pc, return_index=True)
        mate_c[uniq] = pr[idx]; mate_r[pr[idx]] = uniq
    return mate_r, mate_c

def ssp_phases(C, u, v, mate_r, mate_c, verbose=True):
    n = C.shape[0]
    stats = dict(phases=0, levels=0, rows_scanned=0, augment=0)
    while True:
        free_r = np.nonzero(mate_r < 0)[0]
        if len(free_r) == 0: break
        stats['phases'] += 1
        dist = np.full(n, INF, np.int64); pred = -np.ones(n, np.int64)
        settled = np.zeros(n, bool)
        d_row = np.full(n, INF, np.int64); root = -np.ones(n, np.int64)
        d_row[free_r] = 0; root[free_r] = free_r
        frontier = free_r
        reached_rows = [free_r]
        sinks = []
        Dstar = None
        while True:
            # scan frontier rows
            stats['rows_scanned'] += len(frontier)
            R = C[frontier].astype(np.int64) - u[frontier, None] - v[None, :] + d_row[frontier, None]
            R[:, settled] = INF
            best = R.min(0); arg = R.argmin(0)
            upd = best < dist
            dist[upd] = best[upd]; pred[upd] = frontier[arg[upd]]
            # next level
            cand = np.where(settled, INF, dist)
            delta = cand.min()
            if delta >= INF: raise RuntimeError('infeasible')
            stats['levels'] += 1
            newc = np.nonzero(cand == delta)[0]
            settled[newc] = True
            free_cols = newc[mate_c[newc] < 0]
            if len(free_cols):
                Dstar = delta; sinks = free_cols; 
                # rows matched to other newly settled cols are at level Dstar too; they do not need scanning
                break
            frontier = mate_c[newc]
            d_row[frontier] = delta
            root[frontier] = root[pred[newc]]
            reached_rows.append(frontier)
        # dual update
        rr = np.concatenate(reached_rows)
        u[rr] += Dstar - d_row[rr]
        sc = np.nonzero(settled)[0]
        v[sc] -= Dstar - dist[sc]
        # augment: one sink per root
        used_root = set()
        for j in sinks:
            rt = root[pred[j]]
            if rt in used_root: continue
            used_root.add(rt)
            # flip path
            cj = j
            while True:
                i = pred[cj]
                nxt = mate_r[i]
                mate_r[i] = cj; mate_c[cj] = i
                if nxt < 0: break
                cj = nxt
            stats['augment'] += 1
        if verbose and stats['phases'] % 20 == 0:
            print('   phase', stats['phases'], 'free', (mate_r < 0).sum(), 'levels', stats['levels'], 'rows', stats['rows_scanned'], flush=True)
    return stats

def run(name, C):
    n = C.shape[0]
    t = time.time()
    u, v = init_reduce(C)
    mr, mc = greedy_tight(C, u, v)
    f0 = int((mr < 0).sum())
    st = ssp_phases(C, u, v, mr, mc, verbose=False)
    obj = int(C[np.arange(n), mr].sum())
    dt = time.time() - t
    ref = assign_ref.solve_scipy(C)[0]
    print(f"{name}: n={n} free_after_init={f0} phases={st['phases']} levels={st['levels']} rows_scanned={st['rows_scanned']} ({st['rows_scanned']/n:.1f} sweeps) obj={obj} ref={ref} ok={obj==ref} t={dt:.1f}s", flush=True)

if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'small'
    if which == 'small':
        run('1a', g.config1a())
        dist, cab_to, cust_from = g.config1b(); n, c = cost_ref.calculate_cost_np(dist, cab_to, cust_from); run('1b', c)
        run('5a-1000', g.config5a(1000)); run('5b-1000', g.config5b_cost(1000, 200))
        run('2', g.config2()); run('2stand', g.config2_stand())
    elif which == 'mid2':
        run('5a-5000', g.config5a(5000)); run('5b-5000', g.config5b_cost(5000, 1000)); run('2stand', g.config2_stand())
    elif which == 'mid':
        run('5a-5000', g.config5a(5000)); run('5b-5000', g.config5b_cost(5000, 1000))
    elif which == 'big':
        run('5a-20000', g.config5a(20000))
    elif which == 'bigb':
        run('5b-20000', g.config5b_cost(20000, 4000))
