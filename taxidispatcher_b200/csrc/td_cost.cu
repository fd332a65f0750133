// td_cost.cu -- K1: cab x customer cost matrix build.
//
// Replaces calculate_cost() of the reference (split.py:123-136, greedy_opt.py:86-99; the
// `< DROP_TIME` cutoff variant simulate.py:17-33 / Simulator.java:493-520; fill n*n of
// procedure.py:6-12).  cost[i][j] = dist[cab_to[i]][cust_from[j]], square-padded with `fill`.
//
// HBM-write-bound: 4*n^2 bytes out, 4*(n_cabs+n_cust) + 4*S^2 bytes in.  Layout decisions:
//   * one CTA works on one cab row at a time (persistent grid = multiple of the SM count); the
//     stand row dist[cab_to[i]][:] is staged in shared memory once per row, so the per-cell
//     gather hits 32 banks instead of L1 lines;
//   * cust_from[] is staged in shared memory once per CTA (n_cust * 4 B);
//   * each thread emits 4 consecutive customers as one 16-byte streaming store (st.global.cs);
//     rows whose start is not 16-byte aligned (n % 4 != 0) get a scalar head/tail;
//   * cutoff and padding are fused into the same pass.
#include "td_common.cuh"

namespace td {

constexpr int kCostThreads = 512;

template <bool kRowInSmem>
__global__ void __launch_bounds__(kCostThreads)
cost_matrix_kernel(const int32_t *__restrict__ dist, int n_stands,
                   const int32_t *__restrict__ cab_to, int n_cabs,
                   const int32_t *__restrict__ cust_from, int n_cust,
                   int32_t fill, int32_t cutoff, int32_t *__restrict__ cost, int n, int cust_in_smem) {
    extern __shared__ int32_t smem[];
    int32_t *s_row = smem;                                   // n_stands (only if kRowInSmem)
    int32_t *s_cust = smem + (kRowInSmem ? n_stands : 0);    // n_cust   (only if cust_in_smem)
    const int tid = threadIdx.x;
    if (cust_in_smem)
        for (int j = tid; j < n_cust; j += kCostThreads) s_cust[j] = cust_from[j];
    const int32_t *custp = cust_in_smem ? s_cust : cust_from;
    const bool has_cut = cutoff >= 0;

    for (int row = blockIdx.x; row < n; row += gridDim.x) {
        int32_t *out = cost + size_t(row) * n;
        const bool real_row = row < n_cabs;
        const int32_t *drow = nullptr;
        __syncthreads();  // previous row's readers are done with s_row; also orders the s_cust fill
        if (real_row) {
            drow = dist + size_t(cab_to[row]) * n_stands;
            if (kRowInSmem) {
                for (int s = tid; s < n_stands; s += kCostThreads) s_row[s] = drow[s];
                __syncthreads();
                drow = s_row;
            }
        }
        auto cell = [&](int j) -> int32_t {
            if (!real_row || j >= n_cust) return fill;
            int32_t d = drow[custp[j]];
            return (has_cut && d >= cutoff) ? fill : d;
        };
        // scalar head up to the first 16-byte aligned element of this row
        const int head = int((4 - ((size_t(row) * n) & 3)) & 3);
        const int h = head < n ? head : n;
        if (tid < h) out[tid] = cell(tid);
        const int nvec = (n - h) >> 2;
        int4 *outv = reinterpret_cast<int4 *>(out + h);
        for (int v = tid; v < nvec; v += kCostThreads) {
            const int j = h + (v << 2);
            int4 r;
            r.x = cell(j); r.y = cell(j + 1); r.z = cell(j + 2); r.w = cell(j + 3);
            __stcs(outv + v, r);
        }
        const int tail0 = h + (nvec << 2);
        if (tail0 + tid < n && tid < 4) out[tail0 + tid] = cell(tail0 + tid);
    }
}

}  // namespace td

extern "C" int td_cost_matrix(const int32_t *dist, int n_stands, const int32_t *cab_to, int n_cabs,
                              const int32_t *cust_from, int n_cust, int32_t fill, int32_t cutoff,
                              int32_t *cost_out, void *stream) {
    if (n_cabs < 0 || n_cust < 0 || n_stands < 0) return TD_ERR_INVALID;
    const int n = n_cabs > n_cust ? n_cabs : n_cust;
    if (n == 0) return TD_OK;  // simulate.py:21
    if (!cost_out || (n_cabs > 0 && n_cust > 0 && (!dist || !cab_to || !cust_from || n_stands == 0))) return TD_ERR_INVALID;
    if (!td::have_device()) return TD_ERR_NO_DEVICE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t kSmemBudget = 200 * 1024;
    const bool row_in_smem = size_t(n_stands) * 4 <= 96 * 1024;
    size_t smem = row_in_smem ? size_t(n_stands) * 4 : 0;
    const int cust_in_smem = (smem + size_t(n_cust) * 4 <= kSmemBudget) ? 1 : 0;
    if (cust_in_smem) smem += size_t(n_cust) * 4;
    const int sms = td::device_sm_count();
    // enough CTAs per SM to hide the per-row staging latency; never more CTAs than rows
    const size_t per_cta = smem > 1024 ? smem : 1024;
    int per_sm = int((220 * 1024) / per_cta);
    per_sm = per_sm > 4 ? 4 : (per_sm < 1 ? 1 : per_sm);
    int grid = sms * per_sm;
    if (grid > n) grid = n;
    td::ProfScope prof(TD_PROF_COST, st);
    if (row_in_smem) {
        TD_CUDA_TRY(cudaFuncSetAttribute(td::cost_matrix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
        td::cost_matrix_kernel<true><<<grid, td::kCostThreads, smem, st>>>(dist, n_stands, cab_to, n_cabs, cust_from, n_cust,
                                                                           fill, cutoff, cost_out, n, cust_in_smem);
    } else {
        TD_CUDA_TRY(cudaFuncSetAttribute(td::cost_matrix_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
        td::cost_matrix_kernel<false><<<grid, td::kCostThreads, smem, st>>>(dist, n_stands, cab_to, n_cabs, cust_from, n_cust,
                                                                            fill, cutoff, cost_out, n, cust_in_smem);
    }
    TD_LAUNCH_CHECK();
    return TD_OK;
}
