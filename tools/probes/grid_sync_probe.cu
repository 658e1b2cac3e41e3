// Probe: cost of a cooperative launch and of cg::grid.sync() on this GPU, vs plain launches.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_plain(float* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f; }
template <int NS> __global__ void k_coop(float* p) {
    cg::grid_group g = cg::this_grid();
#pragma unroll
    for (int i = 0; i < NS; ++i) g.sync();
    if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}
// hand-rolled barrier: one release-add per CTA, acquire-poll by thread 0
__device__ __forceinline__ void my_sync(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*(volatile unsigned*)ctr < target) {}
        __threadfence();
    }
    __syncthreads();
}
template <int NS> __global__ void k_mine(float* p, unsigned* ctr, unsigned base) {
#pragma unroll
    for (int i = 0; i < NS; ++i) my_sync(ctr, base + (i + 1) * gridDim.x);
    if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}

template <typename F> float time_us(F f, int iters = 200) {
    for (int i = 0; i < 10; ++i) f(i);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f(10 + i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / iters;
}

int main() {
    float* p; cudaMalloc(&p, 4); cudaMemset(p, 0, 4);
    unsigned* ctr; cudaMalloc(&ctr, 4);
    for (int grid : {148, 296, 444, 592}) {
        float t0 = time_us([&](int) { k_plain<<<grid, 256>>>(p); });
        void* args[] = {&p};
        float c0 = time_us([&](int) { cudaLaunchCooperativeKernel((void*)k_coop<0>, dim3(grid), dim3(256), args, 0, 0); });
        float c1 = time_us([&](int) { cudaLaunchCooperativeKernel((void*)k_coop<1>, dim3(grid), dim3(256), args, 0, 0); });
        float c2 = time_us([&](int) { cudaLaunchCooperativeKernel((void*)k_coop<2>, dim3(grid), dim3(256), args, 0, 0); });
        float c8 = time_us([&](int) { cudaLaunchCooperativeKernel((void*)k_coop<8>, dim3(grid), dim3(256), args, 0, 0); });
        cudaMemset(ctr, 0, 4);
        unsigned base = 0;
        float m2 = time_us([&](int) { k_mine<2><<<grid, 256>>>(p, ctr, base); base += 2 * grid; });
        cudaMemset(ctr, 0, 4); cudaDeviceSynchronize(); base = 0;
        float m8 = time_us([&](int) { k_mine<8><<<grid, 256>>>(p, ctr, base); base += 8 * grid; });
        printf("grid %3d: plain %.2f us | coop launch %.2f, +1 sync %.2f, +2 sync %.2f, +8 sync %.2f (%.2f us/sync) | "
               "hand-rolled 2 sync %.2f, 8 sync %.2f (%.2f us/sync)\n",
               grid, t0, c0, c1, c2, c8, (c8 - c2) / 6, m2, m8, (m8 - m2) / 6);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
