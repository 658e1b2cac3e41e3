// Probe for a halo-tile tcgen05 weight gradient: can an MN-MAJOR swizzled operand (rows = K = pixels, columns = M =
// channels, exactly as a channels-last tile lies in shared memory after TMA)
//   (1) start at a pixel row that is not aligned to the 8-row swizzle atom,
//   (2) have its 8-row K groups further apart than 8 rows (SBO = one halo-box row), and
//   (3) have its M "atoms" (the next 32/64 channels) OVERLAP the previous atom shifted by one pixel row (LBO = one row),
//       so that one M=128 instruction stacks the kx = 0,1,2,(3) taps of one kernel row on the M axis?
//
//   S: 256 rows x C channels (bf16 integers), TMA-loaded with the swizzle that matches C (C = 32 -> 64 B rows).
//   A[k][m] = S[shift + (k/8)*group_rows + k%8 + (m/C)*lbo_rows][m % C]     k < 16, m < 128
//   B[k][n] = (k == n), MN-major, so D[m][n] = A[n][m].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn_shift_probe umma_mn_shift_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../vae-2_b200/csrc/tc_ptx.cuh"

using namespace vae2::tc;

struct P { int C, shift, group_rows, lbo_rows; float* out; };

// MN-major descriptor: start, LBO (bytes between M atoms), SBO (bytes between 8-row K groups), swizzle by row width
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, P p) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t row_bytes = p.C * 2;
    uint8_t* sa = smem;                            // 256 rows
    uint8_t* sb = smem + 256 * 128;                // 16 rows x 32 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 256 * row_bytes + 16 * 32);
        tma_load_3d(sa, &map_a, &bars[0], 0, 0, 0);
        tma_load_3d(sb, &map_b, &bars[0], 0, 0, 0);
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        // D=f32, A=B=bf16, A and B MN-major (bits 15, 16), N = 16, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sa) + p.shift * row_bytes;
        umma_bf16(tmem, desc_mn(a0, row_bytes, p.lbo_rows * row_bytes, p.group_rows * row_bytes),
                  desc_mn(smem_u32(sb), 32, 16 * 32, 8 * 32), idesc, 0u);
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
    for (int C : {32, 64, 16}) {
        std::vector<__nv_bfloat16> hA(256 * C), hB(16 * 16);
        std::vector<float> fA(256 * C);
        for (int r = 0; r < 256; ++r)
            for (int c = 0; c < C; ++c) { float v = (float)((r * 7 + c * 3) % 61 - 30); fA[r * C + c] = v; hA[r * C + c] = __float2bfloat16(v); }
        for (int k = 0; k < 16; ++k)
            for (int n = 0; n < 16; ++n) hB[k * 16 + n] = __float2bfloat16(k == n ? 1.f : 0.f);
        __nv_bfloat16 *dA, *dB;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUtensorMap ma, mb;
        {
            cuuint64_t dims[3] = {(cuuint64_t)C, 256, 1}; cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 256};
            cuuint32_t box[3] = {(cuuint32_t)C, 256, 1}; cuuint32_t es[3] = {1, 1, 1};
            if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
        }
        {
            cuuint64_t dims[3] = {16, 16, 1}; cuuint64_t str[2] = {32, 32 * 16};
            cuuint32_t box[3] = {16, 16, 1}; cuuint32_t es[3] = {1, 1, 1};
            if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
        }
        const int atoms = 128 / C;      // M atoms of C channels covered by one M=128 instruction
        for (int lbo_rows : {64, 1})    // 64: atoms from separate tiles (what wgrad_tc_kernel does today); 1: next pixel row (halo)
            for (int group_rows : {8, 10})
                for (int shift = 0; shift < 4; ++shift) {
                    P p{C, shift, group_rows, lbo_rows, d_out};
                    cudaMemset(d_out, 0, 128 * 16 * 4);
                    probe<<<1, 128, 48 * 1024>>>(ma, mb, p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("C=%d: CUDA error %s\n", C, cudaGetErrorString(e)); return 2; }
                    std::vector<float> out(128 * 16);
                    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
                    int bad = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 16; ++n) {
                            const int row = shift + (n / 8) * group_rows + n % 8 + (m / C) * lbo_rows;
                            const float ref = row < 256 ? fA[row * C + m % C] : 0.f;
                            if (out[m * 16 + n] != ref) ++bad;
                        }
                    printf("C=%2d atoms=%d lbo_rows=%2d group_rows=%2d shift=%d : %s (%d / 2048 wrong)\n", C, atoms, lbo_rows, group_rows,
                           shift, bad ? "MISMATCH" : "exact", bad);
                }
        cudaFree(dA); cudaFree(dB);
    }
    return 0;
}
