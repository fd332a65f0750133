// td_cost.cu -- K1: cab x customer cost matrix build.
//
// Replaces calculate_cost() of the reference (split.py:123-136, greedy_opt.py:86-99; the
// `< DROP_TIME` cutoff variant simulate.py:17-33 / Simulator.java:493-520; fill n*n of
// procedure.py:6-12).  cost[i][j] = dist[cab_to[i]][cust_from[j]], square-padded with `fill`.
//
// HBM-write-bound: 4*n^2 bytes out, 4*(n_cabs+n_cust) + 4*S^2 bytes in.  Layout decisions:
//   * persistent grid (a multiple of the SM count); one CTA works on one cab row at a time;
//   * the stand row dist[cab_to[i]][:] is staged in shared memory so the per-cell gather hits 32
//     banks instead of L1 lines; the NEXT row's stand row is prefetched with cp.async into the
//     other half of a double buffer while the current row is being written (one __syncthreads per
//     row, load latency hidden);
//   * cust_from[] is read with 16-byte read-only loads (80 KB at n = 20k: L1-resident);
//   * each thread emits 4 consecutive customers as one 16-byte streaming store (st.global.cs);
//     rows whose start is not 16-byte aligned (n % 4 != 0) get a scalar head/tail;
//   * cutoff and padding are fused into the same pass.
//
// Large matrices take the GROUPED path (td_cost_matrix_rows with a workspace): cab rows that share a stand are identical
// (cost[i][:] depends on cab_to[i] only; 20 000 cabs over 4000 stands: 5 rows per stand), so the rows are bucketed by
// stand first (cost_group_kernel, a counting sort) and every distinct row is gathered ONCE into registers and stored to
// all the rows of its group.  ncu on the row-at-a-time kernel showed the L1/shared-memory data pipe as its limit: the
// gather from the staged stand row costs ~3.5 wavefronts per warp instruction (random banks), 14 of the 22 wavefronts
// per 128 cells; grouped, the kernel is a pure streaming store (measured fill bandwidth of the device: 7.3 TB/s).
//
// Row ranges: [row_begin, row_begin + row_count) of the padded matrix, written to a row_count x n buffer -- the
// multi-GPU path builds contiguous cab-row blocks per rank (north_star: "cost-matrix rows are split
// across devices").
#include "td_common.cuh"
#include <stdlib.h>

namespace td {

constexpr int kCostThreads = 256;
__device__ int g_k1_store_mode = 0;   // EXPERIMENT
__device__ __forceinline__ void st_out(int4 *p, int4 v, int mode) { if (mode == 0) __stcs(p, v); else if (mode == 1) *p = v; else __stwt(p, v); }

__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool kRowInSmem>
__global__ void __launch_bounds__(kCostThreads)
cost_matrix_kernel(const int32_t *__restrict__ dist, int n_stands,
                   const int32_t *__restrict__ cab_to, int n_cabs,
                   const int32_t *__restrict__ cust_from, int n_cust,
                   int32_t fill, int32_t cutoff, int32_t *__restrict__ cost, int n, int row_stride,
                   int row0, int row_end) {   // rows [row0, row_end) of the matrix; cost points at row row0
    extern __shared__ __align__(16) int32_t smem[];  // [2][row_stride] when kRowInSmem
    const int tid = threadIdx.x;
    const int smode = g_k1_store_mode;
    const bool has_cut = cutoff >= 0;
    const bool vec_row = (n_stands & 3) == 0 && ((reinterpret_cast<uintptr_t>(dist) & 15) == 0);
    const bool vec_cust = (reinterpret_cast<uintptr_t>(cust_from) & 15) == 0;

    auto prefetch = [&](int row, int buf) {
        if (!kRowInSmem || row >= n_cabs) return;
        const int32_t *src = dist + size_t(cab_to[row]) * n_stands;
        int32_t *dst = smem + buf * row_stride;
        if (vec_row) for (int s = tid * 4; s < n_stands; s += kCostThreads * 4) cp_async_16(dst + s, src + s);
        else for (int s = tid; s < n_stands; s += kCostThreads) cp_async_4(dst + s, src + s);
    };

    int row = row0 + blockIdx.x;
    prefetch(row, 0);
    for (int it = 0; row < row_end; row += gridDim.x, ++it) {
        const int cur = it & 1;
        cp_async_wait_all();
        __syncthreads();                       // current stand row landed; previous row's readers are done
        prefetch(row + gridDim.x, cur ^ 1);    // overlaps with the stores below
        int32_t *out = cost + size_t(row - row0) * n;
        const bool real_row = row < n_cabs;
        const int32_t *drow = !real_row ? nullptr
                              : (kRowInSmem ? smem + cur * row_stride : dist + size_t(cab_to[row]) * n_stands);
        auto lookup = [&](int stand) -> int32_t {
            const int32_t d = kRowInSmem ? drow[stand] : __ldg(drow + stand);
            return (has_cut && d >= cutoff) ? fill : d;
        };
        auto cell = [&](int j) -> int32_t {
            if (!real_row || j >= n_cust) return fill;
            return lookup(__ldg(cust_from + j));
        };
        // scalar head up to the first 16-byte aligned element of this row
        const int head = int((4 - ((size_t(row - row0) * n) & 3)) & 3);
        const int h = head < n ? head : n;
        if (tid < h) out[tid] = cell(tid);
        const int nvec = (n - h) >> 2;
        int4 *outv = reinterpret_cast<int4 *>(out + h);
        if (!real_row) {
            const int4 f = make_int4(fill, fill, fill, fill);
            for (int v = tid; v < nvec; v += kCostThreads) st_out(outv + v, f, smode);
        } else if (h == 0 && vec_cust) {
#pragma unroll 2
            for (int v = tid; v < nvec; v += kCostThreads) {
                const int j = v << 2;
                int4 r;
                if (j + 3 < n_cust) {
                    const int4 cf = __ldg(reinterpret_cast<const int4 *>(cust_from + j));
                    r.x = lookup(cf.x); r.y = lookup(cf.y); r.z = lookup(cf.z); r.w = lookup(cf.w);
                } else {
                    r.x = cell(j); r.y = cell(j + 1); r.z = cell(j + 2); r.w = cell(j + 3);
                }
                st_out(outv + v, r, smode);
            }
        } else {
            for (int v = tid; v < nvec; v += kCostThreads) {
                const int j = h + (v << 2);
                int4 r;
                r.x = cell(j); r.y = cell(j + 1); r.z = cell(j + 2); r.w = cell(j + 3);
                __stcs(outv + v, r);
            }
        }
        const int tail0 = h + (nvec << 2);
        if (tid < 4 && tail0 + tid < n) out[tail0 + tid] = cell(tail0 + tid);
    }
    cp_async_wait_all();
}


// ---- grouped path ------------------------------------------------------------------------------------------------
// workspace layout (int32): offs[S + 1] | cursor[S] | order[rows] | counter[1]
struct CostGroupWs { int32_t *offs, *cursor, *order, *counter; size_t bytes; };
static CostGroupWs carve_cost(void *ws, int n_stands, int rows) {
    Carver c(ws);
    CostGroupWs w;
    w.offs = c.take<int32_t>(size_t(n_stands) + 1);
    w.cursor = c.take<int32_t>(size_t(n_stands) > 0 ? n_stands : 1);
    w.order = c.take<int32_t>(rows > 0 ? rows : 1);
    w.counter = c.take<int32_t>(4);
    w.bytes = c.used();
    return w;
}

// counting sort of the real rows [r0, r1) by their cab's stand: offs[s] .. offs[s+1] index the rows of stand s in order[].
// One CTA; the histogram and the cursors live in shared memory when the stand count allows (kGroupSmemStands), so the
// two passes over the rows cost shared-memory atomics instead of serialised global ones.
constexpr int kGroupSmemStands = 8192;
template <bool kSmem>
__global__ void __launch_bounds__(1024)
cost_group_kernel(const int32_t *__restrict__ cab_to, int r0, int r1, int n_stands, int32_t *offs, int32_t *cursor,
                  int32_t *order, int32_t *counter) {
    __shared__ int s_part[1024];
    __shared__ int s_hist[kSmem ? kGroupSmemStands + 1 : 1];
    int *hist = kSmem ? s_hist : offs;
    int *cur = kSmem ? s_hist : cursor;      // shared memory: the scanned histogram doubles as the cursor array
    const int tid = threadIdx.x;
    for (int s = tid; s <= n_stands; s += 1024) hist[s] = 0;
    if (!kSmem) for (int s = tid; s < n_stands; s += 1024) cursor[s] = 0;
    if (tid == 0) counter[0] = 0;
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += 1024) atomicAdd(&hist[cab_to[r]], 1);
    __syncthreads();
    // exclusive scan over the stands: per-thread chunks, then the chunk sums
    const int chunk = (n_stands + 1023) / 1024;
    const int lo = min(tid * chunk, n_stands), hi = min(lo + chunk, n_stands);
    int sum = 0;
    for (int s = lo; s < hi; ++s) sum += hist[s];
    s_part[tid] = sum;
    __syncthreads();
    if (tid < 32) {   // warp scan of the 1024 chunk sums (32 per lane)
        int run = 0;
        int loc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { loc[k] = run; run += s_part[tid * 32 + k]; }
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += v; }
        const int excl = incl - run;
#pragma unroll
        for (int k = 0; k < 32; ++k) s_part[tid * 32 + k] = loc[k] + excl;
        if (tid == 31) offs[n_stands] = incl;
    }
    __syncthreads();
    int run = s_part[tid];
    for (int s = lo; s < hi; ++s) { const int v = hist[s]; offs[s] = run; if (kSmem) s_hist[s] = run; run += v; }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += 1024) {
        const int s = cab_to[r];
        const int pos = kSmem ? atomicAdd(&cur[s], 1) : offs[s] + atomicAdd(&cur[s], 1);
        order[pos] = r;
    }
}

// One work item = one stand with at least one row (items 0 .. S-1, empty stands are skipped) or one padding row (items
// S ..).  Items are handed out by an atomic counter; the next item's stand row is prefetched (cp.async) while the
// current one is stored.  Requires n % 4 == 0 and 16-byte aligned cust_from / cost (the caller checks).
constexpr int kGroupRows = 32;   // rows of a group stored per pass over the columns
__global__ void __launch_bounds__(kCostThreads)
cost_matrix_grouped_kernel(const int32_t *__restrict__ dist, int n_stands, const int32_t *__restrict__ cust_from, int n_cust,
                           int32_t fill, int32_t cutoff, int32_t *__restrict__ cost, int n, int row_stride, int row0,
                           int n_fill_rows, int fill_row0, const int32_t *__restrict__ offs, const int32_t *__restrict__ order,
                           int32_t *counter) {
    extern __shared__ __align__(16) int32_t smem[];  // [2][row_stride]
    __shared__ int s_item[2];
    __shared__ int s_rows[kGroupRows];
    const int tid = threadIdx.x;
    const bool has_cut = cutoff >= 0;
    const bool vec_row = (n_stands & 3) == 0 && ((reinterpret_cast<uintptr_t>(dist) & 15) == 0);
    const int n_items = n_stands + n_fill_rows;
    const int nvec = n >> 2;
    auto fetch = [&](int slot) {   // thread 0 takes the next non-empty item
        if (tid == 0) {
            int it;
            do { it = atomicAdd(counter, 1); } while (it < n_stands && offs[it] == offs[it + 1]);
            s_item[slot] = it;
        }
    };
    auto prefetch = [&](int item, int buf) {
        if (item >= n_stands) return;
        const int32_t *src = dist + size_t(item) * n_stands;
        int32_t *dst = smem + buf * row_stride;
        if (vec_row) for (int s = tid * 4; s < n_stands; s += kCostThreads * 4) cp_async_16(dst + s, src + s);
        else for (int s = tid; s < n_stands; s += kCostThreads) cp_async_4(dst + s, src + s);
    };
    fetch(0);
    __syncthreads();
    int item = s_item[0];
    prefetch(item, 0);
    for (int itn = 0; item < n_items; ++itn) {
        const int cur = itn & 1;
        fetch(cur ^ 1);
        cp_async_wait_all();
        __syncthreads();                       // stand row of `item` landed, next item known, previous readers done
        const int next = s_item[cur ^ 1];
        prefetch(next, cur ^ 1);               // overlaps with the stores below
        if (item >= n_stands) {                // a padding row: constant fill
            int4 *outv = reinterpret_cast<int4 *>(cost + size_t(fill_row0 + (item - n_stands) - row0) * n);
            const int4 f = make_int4(fill, fill, fill, fill);
            for (int v = tid; v < nvec; v += kCostThreads) __stcs(outv + v, f);
        } else {
            const int32_t *drow = smem + cur * row_stride;
            auto lookup = [&](int stand) -> int32_t {
                const int32_t d = drow[stand];
                return (has_cut && d >= cutoff) ? fill : d;
            };
            const int g0 = offs[item], g1 = offs[item + 1];
            for (int gb = g0; gb < g1; gb += kGroupRows) {
                const int gn = min(kGroupRows, g1 - gb);
                __syncthreads();
                if (tid < gn) s_rows[tid] = order[gb + tid] - row0;
                __syncthreads();
#pragma unroll 2
                for (int v = tid; v < nvec; v += kCostThreads) {
                    const int j = v << 2;
                    int4 r;
                    if (j + 3 < n_cust) {
                        const int4 cf = __ldg(reinterpret_cast<const int4 *>(cust_from + j));
                        r.x = lookup(cf.x); r.y = lookup(cf.y); r.z = lookup(cf.z); r.w = lookup(cf.w);
                    } else {
                        r.x = j < n_cust ? lookup(__ldg(cust_from + j)) : fill;
                        r.y = j + 1 < n_cust ? lookup(__ldg(cust_from + j + 1)) : fill;
                        r.z = j + 2 < n_cust ? lookup(__ldg(cust_from + j + 2)) : fill;
                        r.w = fill;
                    }
                    for (int g = 0; g < gn; ++g)   // the same 16 bytes go to every row of the group
                        __stcs(reinterpret_cast<int4 *>(cost + size_t(s_rows[g]) * n) + v, r);
                }
            }
        }
        item = next;
    }
    cp_async_wait_all();
}

}  // namespace td

extern "C" size_t td_cost_matrix_workspace_bytes(int n_stands, int row_count) {
    if (n_stands < 0 || row_count < 0) return 0;
    return td::carve_cost(nullptr, n_stands, row_count).bytes;
}

extern "C" int td_cost_matrix_rows(const int32_t *dist, int n_stands, const int32_t *cab_to, int n_cabs,
                                   const int32_t *cust_from, int n_cust, int32_t fill, int32_t cutoff,
                                   int row_begin, int row_count, int32_t *cost_out, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    using namespace td;
    if (n_cabs < 0 || n_cust < 0 || n_stands < 0 || row_begin < 0 || row_count < 0) return TD_ERR_INVALID;
    const int n = n_cabs > n_cust ? n_cabs : n_cust;
    if (row_begin + (long long)row_count > n) return TD_ERR_INVALID;
    if (n == 0 || row_count == 0) return TD_OK;  // simulate.py:21
    if (!cost_out || (n_cabs > 0 && n_cust > 0 && (!dist || !cab_to || !cust_from || n_stands == 0))) return TD_ERR_INVALID;
    if (!have_device()) return TD_ERR_NO_DEVICE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int row_end = row_begin + row_count;
    const int row_stride = (n_stands + 3) & ~3;
    const int sms = device_sm_count();
    const int real_end = row_end < n_cabs ? row_end : n_cabs;          // real rows of the range: [row_begin, real_end)
    const int n_real = real_end > row_begin ? real_end - row_begin : 0;
    // grouped path: pays off when there are rows to share (more real rows than stands is the typical case) and the
    // matrix is big enough to hide the grouping kernel; needs vector-aligned rows
    const bool grouped = workspace && n_cust > 0 && n_real >= 1024 && (long long)n_real * n >= (1ll << 24) && (n & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(cust_from) & 15) == 0 && (reinterpret_cast<uintptr_t>(cost_out) & 15) == 0 &&
                         size_t(row_stride) * 8 <= 96 * 1024 && n_stands <= 2 * n_real &&
                         workspace_bytes >= td_cost_matrix_workspace_bytes(n_stands, row_count);
    ProfScope prof(TD_PROF_COST, st);
    if (grouped) {
        CostGroupWs w = carve_cost(workspace, n_stands, row_count);
        if (n_stands <= kGroupSmemStands)
            cost_group_kernel<true><<<1, 1024, 0, st>>>(cab_to, row_begin, real_end, n_stands, w.offs, w.cursor, w.order, w.counter);
        else
            cost_group_kernel<false><<<1, 1024, 0, st>>>(cab_to, row_begin, real_end, n_stands, w.offs, w.cursor, w.order, w.counter);
        TD_LAUNCH_CHECK();
        const size_t smem = size_t(row_stride) * 8;
        TD_CUDA_TRY(cudaFuncSetAttribute(cost_matrix_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(96 * 1024)));
        int per_sm = 0;
        TD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cost_matrix_grouped_kernel, kCostThreads, smem));
        if (per_sm < 1) return TD_ERR_CUDA;
        per_sm = per_sm > 8 ? 8 : per_sm;
        const int fill_row0 = real_end > row_begin ? real_end : row_begin;
        cost_matrix_grouped_kernel<<<sms * per_sm, kCostThreads, smem, st>>>(dist, n_stands, cust_from, n_cust, fill, cutoff, cost_out,
                                                                             n, row_stride, row_begin, row_end - fill_row0, fill_row0,
                                                                             w.offs, w.order, w.counter);
        TD_LAUNCH_CHECK();
        return TD_OK;
    }
    const bool row_in_smem = size_t(row_stride) * 8 <= 96 * 1024;   // two buffers
    const size_t smem = row_in_smem ? size_t(row_stride) * 8 : 0;
    int per_sm = 8;
    if (smem > 0) { const int fit = int((200 * 1024) / smem); per_sm = fit < per_sm ? fit : per_sm; }
    per_sm = per_sm < 1 ? 1 : per_sm;
    if (const char *e = getenv("TD_K1_PER_SM")) per_sm = atoi(e);   // EXPERIMENT
    if (const char *e = getenv("TD_K1_STORE")) { int m = atoi(e); cudaMemcpyToSymbol(g_k1_store_mode, &m, sizeof m); }
    int grid = sms * per_sm;
    if (grid > row_count) grid = row_count;
    if (row_in_smem) {
        TD_CUDA_TRY(cudaFuncSetAttribute(cost_matrix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(96 * 1024)));
        cost_matrix_kernel<true><<<grid, kCostThreads, smem, st>>>(dist, n_stands, cab_to, n_cabs, cust_from, n_cust, fill, cutoff,
                                                                   cost_out, n, row_stride, row_begin, row_end);
    } else {
        cost_matrix_kernel<false><<<grid, kCostThreads, 0, st>>>(dist, n_stands, cab_to, n_cabs, cust_from, n_cust, fill, cutoff,
                                                                 cost_out, n, row_stride, row_begin, row_end);
    }
    TD_LAUNCH_CHECK();
    return TD_OK;
}

extern "C" int td_cost_matrix(const int32_t *dist, int n_stands, const int32_t *cab_to, int n_cabs,
                              const int32_t *cust_from, int n_cust, int32_t fill, int32_t cutoff,
                              int32_t *cost_out, void *stream) {
    if (n_cabs < 0 || n_cust < 0 || n_stands < 0) return TD_ERR_INVALID;
    const int n = n_cabs > n_cust ? n_cabs : n_cust;
    return td_cost_matrix_rows(dist, n_stands, cab_to, n_cabs, cust_from, n_cust, fill, cutoff, 0, n, cost_out, nullptr, 0, stream);
}
