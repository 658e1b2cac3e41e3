// Probe: can a K-major swizzled tcgen05 A operand start at a row that is NOT aligned to the 8-row swizzle atom,
// and can the 8-row groups be further apart than 8 rows (SBO > 8*row_bytes)?  Both are needed to read the nine
// tap-shifted views of ONE halo tile (3x3 convolution) instead of fetching nine tiles.
//
//   A source: 256 rows x KC channels (bf16 integers), TMA-loaded with the swizzle that matches KC.
//   B: 16 x KC selector, D[m][n] = sum_{c % 16 == n} A[row(m)][c],  row(m) = (m/8)*group_rows + m%8 + shift
//   Variants: shift 0..3, group_rows 8|16, descriptor base_offset 0 | shift | (8-shift)%8.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../vae-2_b200/csrc/tc_ptx.cuh"

using namespace vae2::tc;

struct P { int KC, shift, group_rows, base_off; float* out; };

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes, uint32_t base_off) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, P p) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t row_bytes = p.KC * 2;
    uint8_t* sa = smem;                       // 256 rows
    uint8_t* sb = smem + 256 * row_bytes;     // 16 rows (1024-aligned: 256*32 = 8192 at least)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 256 * row_bytes + 16 * row_bytes);
        tma_load_3d(sa, &map_a, &bars[0], 0, 0, 0);
        tma_load_3d(sb, &map_b, &bars[0], 0, 0, 0);
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sa) + p.shift * row_bytes;
        for (int k = 0; k < p.KC / 16; ++k)
            umma_bf16(tmem, desc_k(a0 + k * 32, row_bytes, p.group_rows * row_bytes, p.base_off),
                      desc_k(smem_u32(sb) + k * 32, row_bytes, 8 * row_bytes, 0), idesc, k ? 1u : 0u);
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    float* d_out; cudaMalloc(&d_out, 128 * 16 * 4);
    for (int KC : {64, 32, 16}) {
        std::vector<__nv_bfloat16> hA(256 * KC), hB(16 * KC);
        std::vector<float> fA(256 * KC);
        for (int r = 0; r < 256; ++r)
            for (int c = 0; c < KC; ++c) { float v = (float)((r * 7 + c * 3) % 17 - 8); fA[r * KC + c] = v; hA[r * KC + c] = __float2bfloat16(v); }
        for (int n = 0; n < 16; ++n)
            for (int c = 0; c < KC; ++c) hB[n * KC + c] = __float2bfloat16(c % 16 == n ? 1.f : 0.f);
        __nv_bfloat16 *dA, *dB;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUtensorMap ma, mb;
        {
            cuuint64_t dims[3] = {(cuuint64_t)KC, 256, 1}; cuuint64_t str[2] = {(cuuint64_t)KC * 2, (cuuint64_t)KC * 2 * 256};
            cuuint32_t box[3] = {(cuuint32_t)KC, 256, 1}; cuuint32_t es[3] = {1, 1, 1};
            if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
        }
        {
            cuuint64_t dims[3] = {(cuuint64_t)KC, 16, 1}; cuuint64_t str[2] = {(cuuint64_t)KC * 2, (cuuint64_t)KC * 2 * 16};
            cuuint32_t box[3] = {(cuuint32_t)KC, 16, 1}; cuuint32_t es[3] = {1, 1, 1};
            if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
        }
        for (int group_rows : {8, 16})
            for (int shift = 0; shift < 4; ++shift) {
                int cands[3] = {0, shift, (8 - shift) % 8};
                for (int ci = 0; ci < 3; ++ci) {
                    if (ci > 0 && cands[ci] == cands[0]) continue;
                    if (ci == 2 && cands[2] == cands[1]) continue;
                    P p{KC, shift, group_rows, cands[ci], d_out};
                    cudaMemset(d_out, 0, 128 * 16 * 4);
                    probe<<<1, 128, 48 * 1024>>>(ma, mb, p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("KC=%d: CUDA error %s\n", KC, cudaGetErrorString(e)); return 2; }
                    std::vector<float> out(128 * 16);
                    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
                    int bad = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 16; ++n) {
                            const int row = (m / 8) * group_rows + m % 8 + shift;
                            float ref = 0.f;
                            for (int c = n; c < KC; c += 16) ref += fA[row * KC + c];
                            if (out[m * 16 + n] != ref) ++bad;
                        }
                    printf("KC=%2d group_rows=%2d shift=%d base_offset=%d : %s (%d / 2048 wrong)\n", KC, group_rows, shift, cands[ci],
                           bad ? "MISMATCH" : "exact", bad);
                }
            }
        cudaFree(dA); cudaFree(dB);
    }
    return 0;
}
