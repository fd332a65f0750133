// Cost of one grid-wide barrier with the geometry of pool_select (2 CTAs of 512 threads per SM): cooperative-groups
// grid.sync() against a hand-rolled ticket barrier.  Build: nvcc -O3 -arch=sm_100a -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 2) k_cg(int iters, unsigned *sink) {
    cg::grid_group grid = cg::this_grid();
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) { acc += i; grid.sync(); }
    if (acc == 0xffffffffu) *sink = acc;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// monotone ticket barrier: every CTA adds 1, waits until the counter reaches the next multiple of the grid size
__device__ __forceinline__ void ticket_barrier(unsigned *counter, unsigned &target, unsigned nblocks) {
    __syncthreads();
    target += nblocks;
    if (threadIdx.x == 0) {
        red_release(counter, 1u);
        while (int(ld_acquire(counter) - target) < 0) {}
    }
    __syncthreads();
}
__global__ void __launch_bounds__(512, 2) k_ticket(int iters, unsigned *counter, unsigned *sink) {
    unsigned target = 0, acc = 0;
    for (int i = 0; i < iters; ++i) { acc += i; ticket_barrier(counter, target, gridDim.x); }
    if (acc == 0xffffffffu) *sink = acc;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned *buf; cudaMalloc(&buf, 256); cudaMemset(buf, 0, 256);
    unsigned *counter = buf, *sink = buf + 32;
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
        dim3 grid(sms * per_sm), block(512);
        void *a1[] = {&iters, &sink};
        void *a2[] = {&iters, &counter, &sink};
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void *)k_cg, grid, block, a1, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("grid %d: cg grid.sync      %.3f us per barrier (%s)\n", grid.x, 1e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
            cudaMemset(counter, 0, 4);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void *)k_ticket, grid, block, a2, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("grid %d: ticket barrier    %.3f us per barrier (%s)\n", grid.x, 1e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
