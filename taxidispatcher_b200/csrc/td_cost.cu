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
// Where the time goes (ncu, 20 000 x 20 000, 4000 stands): 92 % L1/TEX data-pipe utilisation, DRAM at 63 % of its peak,
// 30 % of the issue slots -- the gather from the staged stand row is the limit: 32 lanes hit random banks, about 3.5
// wavefronts per look-up, 14 of the ~22 wavefronts per 128 cells.  The device's fill bandwidth (7.3 TB/s) is therefore
// out of reach of a row-at-a-time gather; the kernel writes at 5.05 TB/s = 0.78 of the measured copy peak.
// Tried and dropped in round 2 (all measured on the B200, none faster than 0.33 ms):
//   * rows bucketed by cab stand, every distinct row gathered once and stored to its whole group (5 rows per stand at
//     20 000 cabs over 4000 stands): 80 % fewer gathers, but the stores of one CTA then go to rows that lie megabytes
//     apart -- 0.34 ms;
//   * 1024-thread CTAs keeping the customer stands of their columns in registers (no global load per row): one CTA per
//     SM cannot hide the latency of the next stand row behind 0.3 us of stores -- 0.37 ms;
//   * fewer CTAs per SM (4: 0.36 ms, 2: 0.51 ms, 1: 0.90 ms), plain / write-through instead of streaming stores (no change).
//
// Row ranges: [row_begin, row_begin + row_count) of the padded matrix, written to a row_count x n buffer -- the
// multi-GPU path builds contiguous cab-row blocks per rank (north_star: "cost-matrix rows are split
// across devices").
#include "td_common.cuh"
#include <stdlib.h>

namespace td {

constexpr int kCostThreads = 256;

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
            for (int v = tid; v < nvec; v += kCostThreads) __stcs(outv + v, f);
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
                __stcs(outv + v, r);
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


}  // namespace td

extern "C" int td_cost_matrix_rows(const int32_t *dist, int n_stands, const int32_t *cab_to, int n_cabs,
                                   const int32_t *cust_from, int n_cust, int32_t fill, int32_t cutoff,
                                   int row_begin, int row_count, int32_t *cost_out, void *stream) {
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
    const bool row_in_smem = size_t(row_stride) * 8 <= 96 * 1024;   // two buffers
    const size_t smem = row_in_smem ? size_t(row_stride) * 8 : 0;
    ProfScope prof(TD_PROF_COST, st);
    int per_sm = 8;
    if (smem > 0) { const int fit = int((200 * 1024) / smem); per_sm = fit < per_sm ? fit : per_sm; }
    per_sm = per_sm < 1 ? 1 : per_sm;
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
    return td_cost_matrix_rows(dist, n_stands, cab_to, n_cabs, cust_from, n_cust, fill, cutoff, 0, n, cost_out, stream);
}
