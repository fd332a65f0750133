/*
 * pool_oracle.c -- CPU restatement of the reference's 2..4-passenger pool finder.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the CHECKER for the CUDA path: it may be
 * called from tests/, from __graft_entry__.smoke() and from bench.py's cpu_baseline
 * leg, never from the product package.  Parity status: PINNED -- validated against the
 * compiled reference (oracle/_ref/pool_n_big, built from /root/reference/pool_n.c by
 * oracle/Makefile) on KAT P1/P2 of SURVEY.md section 4 (see tests/test_oracle.py
 * and tests/golden/pool_*.json).
 *
 * What it restates (reference file:line):
 *   pool_n.c:153-177  findPool       ordered pickup tuples, wait-time pruning
 *   pool_n.c:101-151  drop_customers all drop-off permutations, per-passenger detour test
 *   pool_n.c:187-207  removeDuplicates  sort by cost (stable in practice), greedy disjoint scan
 *   pool_n.c:226-229  shard rule     step = n/8 + 1, [step*t, min(n, step*t+step))
 *   findpool.c:83-108 merge          concatenate shard survivors, sort on column 8, scan again
 *
 * Written from the formal statement in SURVEY.md section 8(a); it is not a copy of the
 * reference source: enumeration is iterative over an explicit odometer, the sort is an
 * explicit stable counting sort on cost, and the disjointness scan uses a per-customer
 * "used" flag (equivalent to the reference's O(P*S) pairwise scan).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define POOL_MAX 4
#define REC_W 9 /* pool_n.c:26 -- 4 pickups + 4 drop-offs + cost at column 8 */

typedef struct {
    int64_t evaluated; /* pool_n.c:103 count_all */
    int64_t feasible;  /* pool_n.c:135 pool_count */
    int64_t kept;      /* pool_n.c:237 good_count */
} pool_oracle_stats;

typedef struct {
    const int32_t *dem; /* n x 5: id, from, to, maxWait, maxLoss  (pool_n.c:20) */
    const int32_t *dist;
    int n, S, k;
    int32_t *rec; /* growing array of feasible records */
    int64_t n_rec, cap_rec;
    int64_t evaluated;
    int oom;
} ctx_t;

#define FROM(c) (x->dem[(c) * 5 + 1])
#define TO(c) (x->dem[(c) * 5 + 2])
#define WAIT(c) (x->dem[(c) * 5 + 3])
#define LOSS(c) (x->dem[(c) * 5 + 4])
#define D(a, b) (x->dist[(int64_t)(a) * x->S + (b)])

static void push_record(ctx_t *x, const int *p, const int *q, int cost) {
    if (x->n_rec == x->cap_rec) {
        int64_t nc = x->cap_rec ? x->cap_rec * 2 : (1 << 16);
        int32_t *nr = (int32_t *)realloc(x->rec, (size_t)nc * REC_W * sizeof(int32_t));
        if (!nr) { x->oom = 1; return; }
        x->rec = nr; x->cap_rec = nc;
    }
    int32_t *r = x->rec + x->n_rec * REC_W;
    memset(r, 0, REC_W * sizeof(int32_t));
    for (int i = 0; i < x->k; i++) { r[i] = p[i]; r[i + x->k] = p[q[i]]; } /* pool_n.c:123-126 */
    r[8] = cost;                                                            /* pool_n.c:134 */
    x->n_rec++;
}

/* All k! drop-off orders of one pickup tuple, lexicographic (pool_n.c:137-150). */
static void eval_tuple(ctx_t *x, const int *p) {
    const int k = x->k;
    int q[POOL_MAX];
    int used[POOL_MAX];
    int lvl = 0;
    memset(used, 0, sizeof used);
    for (int i = 0; i < k; i++) q[i] = -1;
    /* iterative permutation generator in lexicographic order */
    while (lvl >= 0) {
        if (lvl == k) {
            x->evaluated++;
            int happy = 1;
            for (int d = 0; d < k && happy; d++) { /* pool_n.c:105-120 */
                int c = p[q[d]];
                int ride = 0;
                for (int ph = q[d]; ph < k - 1; ph++) ride += D(FROM(p[ph]), FROM(p[ph + 1]));
                ride += D(FROM(p[k - 1]), TO(p[q[0]]));
                for (int ph = 0; ph < d; ph++) ride += D(TO(p[q[ph]]), TO(p[q[ph + 1]]));
                /* int-vs-double compare exactly as the reference writes it (pool_n.c:115-116) */
                if (ride > D(FROM(c), TO(c)) * (1 + LOSS(c) / 100.0)) happy = 0;
            }
            if (happy) {
                int cost = 0; /* pool_n.c:127-134 */
                for (int i = 0; i < k - 1; i++) cost += D(FROM(p[i]), FROM(p[i + 1]));
                cost += D(FROM(p[k - 1]), TO(p[q[0]]));
                for (int i = 0; i < k - 1; i++) cost += D(TO(p[q[i]]), TO(p[q[i + 1]]));
                push_record(x, p, q, cost);
            }
            lvl--;
            continue;
        }
        int c = q[lvl] + 1;
        if (q[lvl] >= 0) used[q[lvl]] = 0;
        while (c < k && used[c]) c++;
        if (c >= k) { q[lvl] = -1; lvl--; continue; }
        q[lvl] = c; used[c] = 1; lvl++;
        if (lvl < k) q[lvl] = -1;
    }
}

/* Ordered pickup tuples with the wait rule (pool_n.c:153-177). */
static void enumerate(ctx_t *x, int start, int stop) {
    const int k = x->k, n = x->n;
    int p[POOL_MAX];
    int lvl = 0;
    p[0] = start - 1;
    while (lvl >= 0) {
        int hi = (lvl == 0) ? stop : n;
        int c = p[lvl] + 1;
        int advanced = 0;
        for (; c < hi; c++) {
            int dup = 0;
            for (int l = 0; l < lvl; l++) if (p[l] == c) { dup = 1; break; }
            if (dup) continue;
            p[lvl] = c;
            int w = 0; /* cumulative pickup distance up to this customer (pool_n.c:169-171) */
            for (int l = 0; l < lvl; l++) w += D(FROM(p[l]), FROM(p[l + 1]));
            if (w > WAIT(c)) continue; /* pool_n.c:172 */
            advanced = 1;
            break;
        }
        if (!advanced) { lvl--; continue; }
        if (lvl == k - 1) {
            eval_tuple(x, p);
        } else {
            lvl++;
            p[lvl] = -1;
        }
    }
}

/* stable sort of records on column 8, then greedy disjoint scan; returns survivors in order */
static int64_t sort_and_select(const int32_t *rec, int64_t n_rec, int n_cust, int width_cmp, int do_sort,
                               int32_t *out, int64_t cap, int64_t *n_out) {
    int64_t *order = (int64_t *)malloc((size_t)(n_rec ? n_rec : 1) * sizeof(int64_t));
    if (!order) return -1;
    if (do_sort && n_rec > 0) {
        int32_t lo = rec[8], hi = rec[8];
        for (int64_t i = 1; i < n_rec; i++) {
            int32_t c = rec[i * REC_W + 8];
            if (c < lo) lo = c;
            if (c > hi) hi = c;
        }
        int64_t span = (int64_t)hi - lo + 1;
        int64_t *cnt = (int64_t *)calloc((size_t)span + 1, sizeof(int64_t));
        if (!cnt) { free(order); return -1; }
        for (int64_t i = 0; i < n_rec; i++) cnt[rec[i * REC_W + 8] - lo + 1]++;
        for (int64_t v = 0; v < span; v++) cnt[v + 1] += cnt[v];
        for (int64_t i = 0; i < n_rec; i++) order[cnt[rec[i * REC_W + 8] - lo]++] = i; /* stable */
        free(cnt);
    } else {
        for (int64_t i = 0; i < n_rec; i++) order[i] = i;
    }
    uint8_t *used = (uint8_t *)calloc((size_t)n_cust + 1, 1);
    if (!used) { free(order); return -1; }
    int64_t kept = 0;
    for (int64_t oi = 0; oi < n_rec; oi++) {
        const int32_t *r = rec + order[oi] * REC_W;
        int clash = 0;
        for (int j = 0; j < width_cmp; j++) if (used[r[j]]) { clash = 1; break; }
        if (clash) continue;
        for (int j = 0; j < width_cmp; j++) used[r[j]] = 1;
        if (kept < cap) memcpy(out + kept * REC_W, r, REC_W * sizeof(int32_t));
        kept++;
    }
    free(used);
    free(order);
    *n_out = kept;
    return 0;
}

/*
 * One logical shard of the search.  dedup != 0: survivors in (cost, enumeration rank) order
 * (what pool_n writes to out<shard>.csv); dedup == 0: every feasible record in enumeration order.
 * Returns 0, or -1 on allocation failure, -2 on bad arguments, -3 when cap is too small
 * (n_plans still holds the required count).
 */
int pool_oracle_find(const int32_t *demand, int n, const int32_t *dist, int n_stands, int pool_size,
                     int shard, int n_shards, int dedup, int32_t *plans_out, int64_t cap,
                     int64_t *n_plans, pool_oracle_stats *st) {
    if (pool_size < 2 || pool_size > POOL_MAX || n < 0 || n_shards < 1 || shard < 0 || shard >= n_shards) return -2;
    ctx_t x;
    memset(&x, 0, sizeof x);
    x.dem = demand; x.dist = dist; x.n = n; x.S = n_stands; x.k = pool_size;
    int step = n / n_shards + 1; /* pool_n.c:226 */
    int start = step * shard;    /* pool_n.c:227 */
    int stop = start + step > n ? n : start + step; /* pool_n.c:228 */
    if (start < n) enumerate(&x, start, stop);
    if (x.oom) { free(x.rec); return -1; }
    int64_t kept = 0;
    int rc = 0;
    if (dedup) {
        /* the reference compares columns 0..3 of each row (pool_n.c:196-197); for k<4 those hold
           the k pickups plus drop-off copies of the same customers, so k columns are equivalent */
        rc = (int)sort_and_select(x.rec, x.n_rec, n, pool_size, 1, plans_out, cap, &kept);
    } else {
        kept = x.n_rec;
        int64_t m = kept < cap ? kept : cap;
        if (m > 0) memcpy(plans_out, x.rec, (size_t)m * REC_W * sizeof(int32_t));
    }
    if (st) { st->evaluated = x.evaluated; st->feasible = x.n_rec; st->kept = kept; }
    *n_plans = kept;
    free(x.rec);
    if (rc) return -1;
    return kept > cap ? -3 : 0;
}

/*
 * findpool.c:83-108 with shard outputs appended in shard-index order (the reference appends in
 * completion order -- SURVEY.md section 4 trap 8).  Reference quirk kept on purpose: findpool's
 * readOutput stores the cost in column 2*pool_size but cmp sorts on column 8 (findpool.c:34-35,70),
 * so for pool_size < 4 the re-sort is a no-op and the scan runs in concatenation order.
 * Rows in/out use the pool_n.c record layout (cost in column 8).
 */
int pool_oracle_merge(const int32_t *plans, int64_t total, int n_cust, int pool_size,
                      int32_t *out, int64_t cap, int64_t *n_out) {
    if (pool_size < 2 || pool_size > POOL_MAX) return -2;
    int64_t kept = 0;
    if (sort_and_select(plans, total, n_cust, pool_size, pool_size == POOL_MAX, out, cap, &kept)) return -1;
    *n_out = kept;
    return kept > cap ? -3 : 0;
}

/*
 * 2-passenger pool of the Simulator: Simulator.java:681-758 (candidates :686-723, TimSort by cost :727,
 * greedy scan :729-739).  accept_all != 0 reproduces `boolean plan1=true, plan2=true` (:691), i.e. what
 * the reference really does (SURVEY.md section 4 trap 6); accept_all == 0 applies the tests as written.
 * out rows: custA, custB, plan (1 = CLNT_B_ENDS, 0 = CLNT_A_ENDS), cost -- in scan order.
 */
int pool_pairs_oracle(const int32_t *from, const int32_t *to, int n, const int32_t *dist, int S, int accept_all,
                      double max_loss, int32_t *out, int64_t cap, int64_t *n_out) {
    int64_t total = (int64_t)n * n, m = 0;
    int32_t *a = (int32_t *)malloc((size_t)(total ? total : 1) * 4 * sizeof(int32_t));
    if (!a) return -1;
#define DD(x, y) dist[(int64_t)(x) * S + (y)]
    for (int A = 0; A < n; A++)
        for (int B = 0; B < n; B++) {
            if (A == B || from[A] < 0 || from[B] < 0) continue;
            int plan1 = accept_all != 0, plan2 = accept_all != 0;
            int cost1 = DD(from[A], from[B]) + DD(from[B], to[A]) + DD(to[A], to[B]);
            int cost2 = DD(from[A], from[B]) + DD(from[B], to[B]) + DD(to[B], to[A]);
            if (DD(from[B], to[A]) + DD(to[A], to[B]) < DD(from[B], to[B]) * max_loss &&
                DD(from[A], from[B]) + DD(from[B], to[A]) < DD(from[A], to[A]) * max_loss) plan1 = 1;
            if (cost2 < DD(from[A], to[A]) * max_loss) plan2 = 1;
            if (plan1 || plan2) {
                int32_t *r = a + m * 4;
                r[0] = A; r[1] = B;
                if (cost1 < cost2) { r[2] = 1; r[3] = cost1; } else { r[2] = 0; r[3] = cost2; }
                m++;
            }
        }
#undef DD
    /* stable counting sort by cost */
    int32_t hi = 0;
    for (int64_t i = 0; i < m; i++) if (a[i * 4 + 3] > hi) hi = a[i * 4 + 3];
    int64_t *cnt = (int64_t *)calloc((size_t)hi + 2, sizeof(int64_t));
    int64_t *ord = (int64_t *)malloc((size_t)(m ? m : 1) * sizeof(int64_t));
    uint8_t *used = (uint8_t *)calloc((size_t)n + 1, 1);
    if (!cnt || !ord || !used) return -1;
    for (int64_t i = 0; i < m; i++) cnt[a[i * 4 + 3] + 1]++;
    for (int32_t v = 0; v <= hi; v++) cnt[v + 1] += cnt[v];
    for (int64_t i = 0; i < m; i++) ord[cnt[a[i * 4 + 3]]++] = i;
    int64_t kept = 0;
    for (int64_t o = 0; o < m; o++) {
        const int32_t *r = a + ord[o] * 4;
        if (used[r[0]] || used[r[1]]) continue;
        used[r[0]] = used[r[1]] = 1;
        if (kept < cap) memcpy(out + kept * 4, r, 4 * sizeof(int32_t));
        kept++;
    }
    free(a); free(cnt); free(ord); free(used);
    *n_out = kept;
    return kept > cap ? -3 : 0;
}
