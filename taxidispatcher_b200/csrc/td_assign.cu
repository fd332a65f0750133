// td_assign.cu -- K2: exact balanced n x n assignment (minimum-cost perfect matching).
//
// Replaces the model build + cvxopt.glpk.ilp call of the reference (solver.py:11-27,
// python.py:6-25, procedure.py:14-29, split.py:139-155, heuristic.py:7-17,37): minimise
// sum c[i][j] x[n*i+j] with every row sum and column sum equal to 1, x binary.  The dense
// 2n x n^2 constraint matrix of solver.py:15-19 (3.5 GB at n = 600) is never materialised.
//
// Method: primal-dual shortest augmenting paths with integer potentials, organised so that every
// step is a bandwidth-bound sweep over whole cost rows and the number of grid-wide steps depends
// on the number of distinct reduced path lengths, not on n:
//
//   init      v_j = min_i c_ij ; u_i = min_j (c_ij - v_j) ; greedy matching on tight cells
//             (each row proposes its first tight free column, lowest row wins; a few rounds).
//             From here on  r_ij = c_ij - u_i - v_j >= 0  and matched cells have r = 0.
//   phase     multi-source Dijkstra from ALL free rows at once.  Costs are integers, so the
//             priority queue is a sequence of levels (Dial): one level = (a) every row that
//             entered the forest at the previous level relaxes all unsettled columns with one
//             coalesced pass over its cost row (warp tiles of 256 columns, 64-bit atomicMin of
//             (distance, predecessor row) per column), (b) all columns at the new minimum distance
//             are settled together; matched ones pull their mate row into the forest, free ones
//             are sinks.  The phase ends at the first level that reaches a sink; potentials are
//             updated (u_i += D - d_i on forest rows, v_j -= D - dist_j on settled columns), which
//             makes every forest edge tight, and one augmenting path per tree is flipped (trees
//             are vertex-disjoint because every column has one predecessor and every matched row
//             one parent column).  Each phase matches at least one more row.
//   finish    objective = sum c[i][mate(i)] in int64; optional dense x in the reference layout.
//
// Exactness: potentials stay feasible and matched cells stay tight through every step, so when
// the matching is perfect, complementary slackness gives optimality for any integer costs
// (ties included) -- no epsilon, no scaling, no price wars on the reference's tiny cost ranges.
//
// Everything runs in ONE cooperative persistent kernel (grid = resident CTAs of all 148 SMs);
// steps are separated by grid.sync().  Algorithmic bytes: 4n bytes per row relaxed; the solver
// reports rows_scanned so achieved GB/s = 4 n rows_scanned / time.
#include "td_common.cuh"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace td {

constexpr int kAsgThreads = 512;
constexpr int kRowBits = 24;                       // predecessor row in the low bits of the packed key
constexpr unsigned long long kDistInf = ~0ull;
constexpr long long kRowInf = (1ll << 62);
constexpr int kGreedyRounds = 4;

struct AsgCtrl {
    unsigned long long gmin[3];
    unsigned int fcount[2];
    unsigned int nsinks;
    unsigned int free_after_init;
    unsigned int phases, levels, augment, pad;
    unsigned long long rows_scanned;
    unsigned long long stale_cells;   // cells re-read to refresh cached level-0 distances
    long long sfree;                  // sum of the phase lengths D: every still-free row has u = u_init + sfree
    unsigned int nstale, pad3;
    unsigned long long t_prof[16];     // ns spent by CTA 0 in: scan, sync, settle, sync, phase start, augment (diagnostics)
    long long objective;
    int status;
    int pad2;
};

struct AsgArgs {
    const int32_t *cost; int n;
    long long *u, *v, *drow, *uinit, *vinit;
    unsigned long long *base0;        // per column: min over the FREE rows of (c - u_init - v_init), packed with the row
    int32_t *stale;
    int32_t *vmin; int32_t *mate_r, *mate_c, *root, *claim, *prop, *argcol;
    unsigned long long *distpred; uint8_t *settled;
    int32_t *frontier[2]; int32_t *sinks;
    AsgCtrl *ctrl;
    int32_t *col_of_row_out; long long *objective_out; uint8_t *x_out;
    int max_phases;
};

__device__ __forceinline__ unsigned long long pack_dp(long long dist, int row) {
    return ((unsigned long long)dist << kRowBits) | (unsigned)row;
}
__device__ __forceinline__ long long dp_dist(unsigned long long k) { return (long long)(k >> kRowBits); }
__device__ __forceinline__ int dp_row(unsigned long long k) { return int(k & ((1u << kRowBits) - 1)); }

// One warp relaxes a 256-column tile for a strided group of rows.  rows == nullptr: rows are 0..nrows-1.
// kMode 0: column minima of the raw costs (init);  kMode 1: Dijkstra relaxation.
template <int kMode, bool kVec>
__device__ __forceinline__ void sweep_rows(const AsgArgs &a, const int32_t *rows, int nrows, int gwarp, int nwarps,
                                           int lane, unsigned long long &block_min) {
    const int n = a.n;
    const int tiles = (n + 255) >> 8;
    int groups = nwarps / tiles;
    groups = groups < 1 ? 1 : (groups > nrows ? nrows : groups);
    const int units = tiles * groups;
    for (int unit = gwarp; unit < units; unit += nwarps) {
        const int tile = unit % tiles, grp = unit / tiles;
        const int j0 = (tile << 8) + (kVec ? lane * 4 : lane);
        // column slots of this lane: kVec -> {j0..j0+3, j0+128..j0+131}; scalar -> j0 + 32*k
        int col[8];
        bool act[8];
        long long vj[8];
        unsigned long long best[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            col[k] = kVec ? j0 + (k & 3) + ((k >> 2) << 7) : j0 + (k << 5);
            act[k] = col[k] < n;
            if (kMode == 1) {
                act[k] = act[k] && !a.settled[act[k] ? col[k] : 0];
                vj[k] = act[k] ? a.v[col[k]] : 0;
            } else {
                vj[k] = 0;
            }
            best[k] = kDistInf;
        }
        // 4 rows per trip: all index / potential / cost loads are issued before any is consumed
        for (int r = grp; r < nrows; r += 4 * groups) {
            int ri[4];
            long long base[4];
            int4 lo[4], hi[4];
            int cs[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rr = r + q * groups;
                ri[q] = rr < nrows ? (rows ? rows[rr] : rr) : -1;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (ri[q] < 0) continue;
                const int32_t *line = a.cost + size_t(ri[q]) * n;
                if (kVec) {
                    lo[q] = (j0 < n) ? ld_stream_int4(reinterpret_cast<const int4 *>(line + j0)) : make_int4(0, 0, 0, 0);
                    hi[q] = (j0 + 128 < n) ? ld_stream_int4(reinterpret_cast<const int4 *>(line + j0 + 128)) : make_int4(0, 0, 0, 0);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) cs[q][k] = col[k] < n ? __ldg(line + col[k]) : 0;
                }
                if (kMode == 1) base[q] = a.drow[ri[q]] - a.u[ri[q]];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (ri[q] < 0) continue;
                int c[8];
                if (kVec) {
                    c[0] = lo[q].x; c[1] = lo[q].y; c[2] = lo[q].z; c[3] = lo[q].w;
                    c[4] = hi[q].x; c[5] = hi[q].y; c[6] = hi[q].z; c[7] = hi[q].w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = cs[q][k];
                }
                if (kMode == 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // order-preserving bias so that negative costs work too
                        const unsigned long long key = (unsigned long long)(unsigned(c[k]) ^ 0x80000000u);
                        best[k] = key < best[k] ? key : best[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const unsigned long long key = pack_dp(base[q] + c[k] - vj[k], ri[q]);
                        best[k] = key < best[k] ? key : best[k];
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!act[k] || best[k] == kDistInf) continue;
            if (kMode == 0) {
                const int32_t val = int32_t(unsigned(best[k]) ^ 0x80000000u);
                if (val < a.vmin[col[k]]) atomicMin(&a.vmin[col[k]], val);
            } else {
                if (best[k] < a.distpred[col[k]]) atomicMin(&a.distpred[col[k]], best[k]);
                block_min = best[k] < block_min ? best[k] : block_min;
            }
        }
    }
}

// warp per row: minimum and first argmin column of (c_ij - v_j); optionally restricted to free, tight columns
// Ties are broken by the column's cyclic distance from a row-dependent offset, so that rows with
// many equally good columns spread their proposals instead of all asking for the same few.
__device__ __forceinline__ unsigned row_offset(int i, int n) { return unsigned((unsigned(i) * 2654435761u) % unsigned(n)); }

template <bool kVec>
__device__ __forceinline__ unsigned long long row_min_reduced(const AsgArgs &a, int i, int lane, bool tight_free_only,
                                                              long long ui) {
    const int n = a.n;
    const unsigned off = row_offset(i, n);
    auto rot = [&](int j) -> unsigned { const unsigned d = unsigned(j) + unsigned(n) - off; return d >= unsigned(n) ? d - unsigned(n) : d; };
    const int32_t *line = a.cost + size_t(i) * n;
    unsigned long long best = kDistInf;
    if (kVec) {
        // 4 x 16-byte loads in flight per lane (a warp-per-row scan is latency-bound otherwise)
        for (int j0 = lane * 4; j0 < n; j0 += 512) {
            int4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 128 * u;
                q[u] = j < n ? ld_stream_int4(reinterpret_cast<const int4 *>(line + j)) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 128 * u;
                if (j >= n) continue;
                const int c[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long red = (long long)c[k] - a.v[j + k];
                    if (tight_free_only) {
                        if (red == ui && a.mate_c[j + k] < 0) { const unsigned long long key = rot(j + k); best = key < best ? key : best; }
                    } else {
                        const unsigned long long key = ((unsigned long long)red << 32) | rot(j + k);
                        best = key < best ? key : best;
                    }
                }
            }
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const long long red = (long long)__ldg(line + j) - a.v[j];
            if (tight_free_only) {
                if (red == ui && a.mate_c[j] < 0) { const unsigned long long key = rot(j); best = key < best ? key : best; }
            } else {
                const unsigned long long key = ((unsigned long long)red << 32) | rot(j);
                best = key < best ? key : best;
            }
        }
    }
    return warp_min_u64(best);
}

template <bool kVec>
__global__ void __launch_bounds__(kAsgThreads)
assign_kernel(AsgArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int n = a.n;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int gwarp = tid >> 5, nwarps = nthreads >> 5;
    __shared__ unsigned long long s_red[kAsgThreads / 32];
    AsgCtrl *ctrl = a.ctrl;
    unsigned long long t_last = 0;
    auto tick = [&](int k) {   // thread 0 only: accumulate wall time since the previous tick into bucket k
        if (tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (k >= 0) ctrl->t_prof[k] += t - t_last;
            t_last = t;
        }
    };

    // ================= init =====================================================================
    for (int j = tid; j < n; j += nthreads) { a.vmin[j] = INT_MAX; a.mate_c[j] = -1; a.prop[j] = INT_MAX; }
    for (int i = tid; i < n; i += nthreads) { a.mate_r[i] = -1; }
    if (tid == 0) {
        ctrl->fcount[0] = ctrl->fcount[1] = 0; ctrl->nsinks = 0; ctrl->phases = ctrl->levels = ctrl->augment = 0;
        ctrl->rows_scanned = 0; ctrl->status = TD_OK; ctrl->gmin[0] = ctrl->gmin[1] = ctrl->gmin[2] = kDistInf;
    }
    grid.sync();
    tick(-1);
    {   // column minima
        unsigned long long dummy = kDistInf;
        sweep_rows<0, kVec>(a, nullptr, n, gwarp, nwarps, lane, dummy);
    }
    grid.sync();
    tick(6);
    for (int j = tid; j < n; j += nthreads) a.v[j] = a.vmin[j];
    grid.sync();
    // row minima of the column-reduced costs; every row proposes its first argmin column
    for (int i = gwarp; i < n; i += nwarps) {
        const unsigned long long k = row_min_reduced<kVec>(a, i, lane, false, 0);
        if (lane == 0) {
            a.u[i] = (long long)(k >> 32);
            const int j = int((unsigned(k) + row_offset(i, n)) % unsigned(n));
            a.argcol[i] = j;
            atomicMin(&a.prop[j], i);
        }
    }
    grid.sync();
    tick(7);
    for (int round = 0;; ++round) {
        // accept: the lowest proposing row takes the column
        for (int i = tid; i < n; i += nthreads) {
            if (a.mate_r[i] >= 0) continue;
            const int j = a.argcol[i];
            if (j >= 0 && a.prop[j] == i) { a.mate_r[i] = j; a.mate_c[j] = i; }
        }
        grid.sync();
        if (round == kGreedyRounds - 1) break;
        for (int j = tid; j < n; j += nthreads) a.prop[j] = INT_MAX;
        grid.sync();
        // losers look for their first tight column that is still free
        for (int i = gwarp; i < n; i += nwarps) {
            if (a.mate_r[i] >= 0) continue;
            const unsigned long long k = row_min_reduced<kVec>(a, i, lane, true, a.u[i]);
            if (lane == 0) {
                const int j = (k == kDistInf) ? -1 : int((unsigned(k) + row_offset(i, n)) % unsigned(n));
                a.argcol[i] = j;
                if (j >= 0) atomicMin(&a.prop[j], i);
            }
        }
        grid.sync();
    }

    for (int i = tid; i < n; i += nthreads) { a.uinit[i] = a.u[i]; a.vinit[i] = a.v[i]; }
    // ================= phases ===================================================================
    bool first_phase = true;
    for (int phase = 0;; ++phase) {
        // ---- P0: reset search state, frontier = all free rows ---------------------------------
        tick(-1);
        for (int j = tid; j < n; j += nthreads) { a.distpred[j] = kDistInf; a.settled[j] = 0; }
        for (int base = blockIdx.x * blockDim.x; base < n; base += nthreads) {
            const int i = base + threadIdx.x;
            bool is_free = false;
            if (i < n) {
                a.claim[i] = INT_MAX;
                is_free = a.mate_r[i] < 0;
                a.drow[i] = is_free ? 0 : kRowInf;
                if (is_free) a.root[i] = i;
            }
            const unsigned ball = __ballot_sync(0xffffffffu, is_free);
            unsigned wb = 0;
            if (lane == 0 && ball) wb = atomicAdd(&ctrl->fcount[0], __popc(ball));
            wb = __shfl_sync(0xffffffffu, wb, 0);
            if (is_free) a.frontier[0][wb + __popc(ball & ((1u << lane) - 1))] = i;
        }
        grid.sync();
        const unsigned nfree = ctrl->fcount[0];
        if (first_phase && tid == 0) ctrl->free_after_init = nfree;
        first_phase = false;
        if (nfree == 0) break;
        if (phase >= a.max_phases) { if (tid == 0) ctrl->status = TD_ERR_NOT_CONVERGED; break; }
        if (phase > 0) {
            // Level 0 without re-reading the free rows.  Every free row has been free since the start and has
            // received the same potential shift (sfree), and rows only LEAVE the free set, so the column minimum
            // over the free rows taken once (base0) stays valid until its arg-min row is matched; only those
            // columns are refreshed (a strided read of the column over the free-row list).
            for (int j = tid; j < n; j += nthreads)
                if (a.mate_r[dp_row(a.base0[j])] >= 0) a.stale[atomicAdd(&ctrl->nstale, 1u)] = j;
            grid.sync();
            const unsigned ns_cols = ctrl->nstale;
            for (unsigned sidx = gwarp; sidx < ns_cols; sidx += nwarps) {
                const int j = a.stale[sidx];
                const long long vj0 = a.vinit[j];
                unsigned long long best = kDistInf;
                for (unsigned t = lane; t < nfree; t += 32) {
                    const int i = a.frontier[0][t];
                    const unsigned long long key = pack_dp((long long)__ldg(a.cost + size_t(i) * n + j) - a.uinit[i] - vj0, i);
                    best = key < best ? key : best;
                }
                best = warp_min_u64(best);
                if (lane == 0) a.base0[j] = best;
            }
            if (tid == 0) ctrl->stale_cells += (unsigned long long)ns_cols * nfree;
            grid.sync();
            const long long sfree = ctrl->sfree;
            unsigned long long lmin = kDistInf;
            for (int j = tid; j < n; j += nthreads) {
                const unsigned long long b = a.base0[j];
                const long long d0 = dp_dist(b) + a.vinit[j] - a.v[j] - sfree;
                a.distpred[j] = pack_dp(d0, dp_row(b));
                lmin = (unsigned long long)d0 < lmin ? (unsigned long long)d0 : lmin;
            }
            lmin = warp_min_u64(lmin);
            if (lane == 0 && lmin != kDistInf) atomicMin(&ctrl->gmin[0], lmin);
            if (tid == 0) ctrl->nstale = 0;
        }

        int cur = 0;
        long long dstar = 0;
        tick(4);
        for (int level = 0;; ++level) {
            const int slot = level % 3;
            // ---- (a) relax: every frontier row against all unsettled columns ------------------
            const unsigned fc = (phase > 0 && level == 0) ? 0u : ctrl->fcount[cur];   // level 0 comes from the cache
            unsigned long long bmin = kDistInf;
            sweep_rows<1, kVec>(a, a.frontier[cur], int(fc), gwarp, nwarps, lane, bmin);
            bmin = warp_min_u64(bmin);
            if (lane == 0) s_red[threadIdx.x >> 5] = bmin;
            __syncthreads();
            if (threadIdx.x < 32) {
                unsigned long long m = threadIdx.x < kAsgThreads / 32 ? s_red[threadIdx.x] : kDistInf;
                m = warp_min_u64(m);
                if (threadIdx.x == 0 && m != kDistInf) atomicMin(&ctrl->gmin[slot], m >> kRowBits);
            }
            if (tid == 0) ctrl->rows_scanned += fc;
            if (tid == 0) {   // diagnostics: levels and scan time by frontier size (<= 32, <= 256, <= 2048, larger)
                const int bkt = fc <= 32 ? 0 : (fc <= 256 ? 1 : (fc <= 2048 ? 2 : 3));
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                ctrl->t_prof[8 + bkt] += t - t_last;
                ctrl->t_prof[12 + bkt] += 1;
            }
            tick(0);
            grid.sync();
            tick(1);
            // ---- (b) settle every column at the new minimum distance -------------------------
            const unsigned long long dl = ctrl->gmin[slot];
            if (dl == kDistInf) { if (tid == 0) ctrl->status = TD_ERR_NOT_CONVERGED; dstar = -1; break; }
            const long long delta = (long long)dl;
            unsigned long long carry = kDistInf;
            for (int base = blockIdx.x * blockDim.x; base < n; base += nthreads) {
                const int j = base + threadIdx.x;
                bool push = false;
                int mate = -1;
                if (phase == 0 && level == 0 && j < n) a.base0[j] = a.distpred[j];   // sfree = 0, v = v_init here
                if (j < n && !a.settled[j]) {
                    const unsigned long long k = a.distpred[j];
                    if (k != kDistInf) {
                        const long long d = dp_dist(k);
                        if (d == delta) {
                            a.settled[j] = 1;
                            mate = a.mate_c[j];
                            if (mate < 0) {
                                a.sinks[atomicAdd(&ctrl->nsinks, 1u)] = j;
                            } else {
                                a.drow[mate] = delta;
                                a.root[mate] = a.root[dp_row(k)];
                                push = true;
                            }
                        } else {
                            carry = (unsigned long long)d < carry ? (unsigned long long)d : carry;
                        }
                    }
                }
                const unsigned ball = __ballot_sync(0xffffffffu, push);
                unsigned wb = 0;
                if (lane == 0 && ball) wb = atomicAdd(&ctrl->fcount[cur ^ 1], __popc(ball));
                wb = __shfl_sync(0xffffffffu, wb, 0);
                if (push) a.frontier[cur ^ 1][wb + __popc(ball & ((1u << lane) - 1))] = mate;
            }
            carry = warp_min_u64(carry);
            if (lane == 0 && carry != kDistInf) atomicMin(&ctrl->gmin[(level + 1) % 3], carry);
            if (tid == 0) { ctrl->gmin[(level + 2) % 3] = kDistInf; ctrl->levels += 1; }
            tick(2);
            grid.sync();
            tick(3);
            if (tid == 0) ctrl->fcount[cur] = 0;  // consumed; becomes the target two levels from now
            if (ctrl->nsinks > 0) { dstar = delta; break; }
            cur ^= 1;
        }
        if (dstar < 0) break;

        // ---- augment: one sink per tree, smallest column index wins ---------------------------
        const unsigned ns = ctrl->nsinks;
        for (unsigned s = tid; s < ns; s += nthreads) {
            const int j = a.sinks[s];
            atomicMin(&a.claim[a.root[dp_row(a.distpred[j])]], j);
        }
        grid.sync();
        for (unsigned s = tid; s < ns; s += nthreads) {
            int j = a.sinks[s];
            if (a.claim[a.root[dp_row(a.distpred[j])]] != j) continue;
            for (;;) {  // flip the tree path sink -> root
                const int i = dp_row(a.distpred[j]);
                const int nxt = a.mate_r[i];
                a.mate_r[i] = j;
                a.mate_c[j] = i;
                if (nxt < 0) break;
                j = nxt;
            }
            atomicAdd(&ctrl->augment, 1u);
        }
        // potentials: every forest edge becomes tight, feasibility is kept
        for (int i = tid; i < n; i += nthreads) {
            const long long d = a.drow[i];
            if (d != kRowInf) a.u[i] += dstar - d;
        }
        for (int j = tid; j < n; j += nthreads)
            if (a.settled[j]) a.v[j] -= dstar - dp_dist(a.distpred[j]);
        if (tid == 0) {
            ctrl->nsinks = 0; ctrl->fcount[0] = ctrl->fcount[1] = 0; ctrl->phases += 1; ctrl->sfree += dstar;
            ctrl->gmin[0] = ctrl->gmin[1] = ctrl->gmin[2] = kDistInf;
        }
        grid.sync();
        tick(5);
    }

    // ================= finish ===================================================================
    long long part = 0;
    for (int i = tid; i < n; i += nthreads) {
        const int j = a.mate_r[i];
        a.col_of_row_out[i] = j;
        if (j >= 0) {
            part += a.cost[size_t(i) * n + j];
            if (a.x_out) a.x_out[size_t(i) * n + j] = 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0 && part != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&ctrl->objective), (unsigned long long)part);
    grid.sync();
    if (tid == 0) *a.objective_out = ctrl->objective;
}

static AsgArgs carve_assign(void *ws, int n, size_t *bytes) {
    Carver c(ws);
    AsgArgs a;
    memset(&a, 0, sizeof a);
    const size_t nn = n > 0 ? n : 1;
    a.ctrl = c.take<AsgCtrl>(1);
    a.u = c.take<long long>(nn); a.v = c.take<long long>(nn); a.drow = c.take<long long>(nn);
    a.uinit = c.take<long long>(nn); a.vinit = c.take<long long>(nn);
    a.base0 = c.take<unsigned long long>(nn); a.stale = c.take<int32_t>(nn);
    a.distpred = c.take<unsigned long long>(nn);
    a.vmin = c.take<int32_t>(nn); a.mate_r = c.take<int32_t>(nn); a.mate_c = c.take<int32_t>(nn);
    a.root = c.take<int32_t>(nn); a.claim = c.take<int32_t>(nn); a.prop = c.take<int32_t>(nn); a.argcol = c.take<int32_t>(nn);
    a.frontier[0] = c.take<int32_t>(nn); a.frontier[1] = c.take<int32_t>(nn); a.sinks = c.take<int32_t>(nn);
    a.settled = c.take<uint8_t>(nn);
    *bytes = c.used();
    return a;
}

}  // namespace td

extern "C" size_t td_assign_workspace_bytes(int n) {
    size_t b = 0;
    td::carve_assign(nullptr, n, &b);
    return b;
}

extern "C" int td_assign_exact(const int32_t *cost, int n, int32_t *col_of_row_out, int64_t *objective_out, uint8_t *x_out,
                               td_assign_stats *stats, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace td;
    if (n < 0 || n >= (1 << kRowBits)) return TD_ERR_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    if (!have_device()) return TD_ERR_NO_DEVICE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (objective_out) TD_CUDA_TRY(cudaMemsetAsync(objective_out, 0, sizeof(int64_t), st));
        return TD_OK;
    }
    if (!cost || !col_of_row_out || !objective_out || !workspace) return TD_ERR_INVALID;
    if (workspace_bytes < td_assign_workspace_bytes(n)) return TD_ERR_WORKSPACE;
    size_t bytes = 0;
    AsgArgs a = carve_assign(workspace, n, &bytes);
    a.cost = cost; a.n = n; a.col_of_row_out = col_of_row_out; a.objective_out = reinterpret_cast<long long *>(objective_out);
    a.x_out = x_out; a.max_phases = n + 8;
    TD_CUDA_TRY(cudaMemsetAsync(a.ctrl, 0, sizeof(AsgCtrl), st));
    if (x_out) TD_CUDA_TRY(cudaMemsetAsync(x_out, 0, size_t(n) * n, st));
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
    void *kern = vec ? (void *)assign_kernel<true> : (void *)assign_kernel<false>;
    int per_sm = 0;
    if (vec) TD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, assign_kernel<true>, kAsgThreads, 0));
    else TD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, assign_kernel<false>, kAsgThreads, 0));
    if (per_sm < 1) return TD_ERR_CUDA;
    per_sm = per_sm > 2 ? 2 : per_sm;
    int grid = device_sm_count() * per_sm;
    // small problems: fewer CTAs make grid.sync cheaper; keep at least one warp per 256-column tile and row
    const long long want_warps = (long long)((n + 255) / 256) * (n < 64 ? n : 64);
    const int need = int((want_warps * 32 + kAsgThreads - 1) / kAsgThreads);
    if (grid > need) grid = need < 1 ? 1 : need;
    void *args[] = {(void *)&a};
    {
        ProfScope prof(TD_PROF_ASSIGN, st);
        TD_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kAsgThreads), args, 0, st));
    }
    count_launch();
    if (stats) {
        AsgCtrl h;
        TD_CUDA_TRY(cudaMemcpyAsync(&h, a.ctrl, sizeof h, cudaMemcpyDeviceToHost, st));
        TD_CUDA_TRY(cudaStreamSynchronize(st));
        stats->objective = h.objective;
        // + the two init sweeps + the strided refreshes of cached level-0 minima (one cost cell each)
        stats->rows_scanned = int64_t(h.rows_scanned) + 2 * int64_t(n) + int64_t(h.stale_cells / (unsigned long long)n);
        stats->auction_rounds = kGreedyRounds;
        stats->phases = int32_t(h.phases);
        stats->search_steps = int32_t(h.levels);
        stats->augmentations = int32_t(h.augment);
        stats->unassigned_after_auction = int32_t(h.free_after_init);
        if (getenv("TD_ASSIGN_PROF"))
            fprintf(stderr, "[td_assign] us: scan %.0f sync1 %.0f settle %.0f sync2 %.0f phase_start %.0f augment %.0f\n",
                    h.t_prof[0] / 1e3, h.t_prof[1] / 1e3, h.t_prof[2] / 1e3, h.t_prof[3] / 1e3, h.t_prof[4] / 1e3, h.t_prof[5] / 1e3);
        if (getenv("TD_ASSIGN_PROF"))
            fprintf(stderr, "[td_assign] init us: column-min sweep %.0f, row-min sweep %.0f\n", h.t_prof[6] / 1e3, h.t_prof[7] / 1e3);
        if (getenv("TD_ASSIGN_PROF"))
            fprintf(stderr, "[td_assign] scan us by frontier size <=32: %.0f (%llu levels)  <=256: %.0f (%llu)  <=2048: %.0f (%llu)  >2048: %.0f (%llu)\n",
                    h.t_prof[8] / 1e3, h.t_prof[12], h.t_prof[9] / 1e3, h.t_prof[13], h.t_prof[10] / 1e3, h.t_prof[14],
                    h.t_prof[11] / 1e3, h.t_prof[15]);
        if (h.status != TD_OK) return h.status;
    }
    return TD_OK;
}
