// td_common.cuh -- shared helpers for the libtaxidispatch kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <limits.h>
#include "../../include/taxidispatch.h"

namespace td {

constexpr int kNumSMsFallback = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing --------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char *what);
void count_launch(int n = 1);
int device_sm_count();
bool have_device();

// RAII: records a start/stop event pair around a launch when profiling is enabled (td_prof_enable)
struct ProfScope {
    int kind; cudaStream_t st; void *slot;
    ProfScope(int kind, cudaStream_t st);
    ~ProfScope();
};

#define TD_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            td::set_cuda_error(_e, #expr);                  \
            return TD_ERR_CUDA;                             \
        }                                                   \
    } while (0)

#define TD_LAUNCH_CHECK()                                   \
    do {                                                    \
        td::count_launch();                                 \
        cudaError_t _e = cudaPeekAtLastError();             \
        if (_e != cudaSuccess) {                            \
            td::set_cuda_error(_e, "kernel launch");        \
            return TD_ERR_CUDA;                             \
        }                                                   \
    } while (0)

// ---- workspace carving (256-byte aligned slices of the caller's buffer) ------------------------
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base(static_cast<char *>(p)), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T *p = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return p;
    }
    size_t used() const { return (off + 255) & ~size_t(255); }
};

// ---- order-preserving keys ---------------------------------------------------------------------
// (value, flat index) packed so that unsigned 64-bit order == (signed value, index) order: this is
// exactly numpy's first-index argmin / Simulator.java:531-537 strict '<' scan.
__host__ __device__ __forceinline__ uint64_t pack_key(int32_t v, uint32_t idx) {
    return (uint64_t(uint32_t(v) ^ 0x80000000u) << 32) | idx;
}
__host__ __device__ __forceinline__ int32_t key_value(uint64_t k) { return int32_t(uint32_t(k >> 32) ^ 0x80000000u); }
__host__ __device__ __forceinline__ uint32_t key_index(uint64_t k) { return uint32_t(k); }
constexpr uint64_t kKeyInf = ~uint64_t(0);

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    return v;
}

__device__ __forceinline__ int4 ld_stream_int4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

}  // namespace td
