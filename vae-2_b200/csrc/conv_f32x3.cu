// fp32-accurate convolution on tcgen05 tensor cores: three-way bf16 split ("bf16x3") with fp32 accumulation.
//
// The fp32 configuration (BASELINE configs[1]) must stay within 1e-4 of the reference.  A single TF32 pass
// (~1e-3 per product) cannot do that, and a 2-term TF32 split only represents 22 of the 24 significand bits
// (measured 4e-6 per conv, 20x the error of an fp32 FMA loop).  Every fp32 operand is therefore split EXACTLY
// into three bf16 terms,  x = x1 + x2 + x3  (8 + 8 + 8 significand bits), and the product is accumulated as
//     x1*w1 + x1*w2 + x2*w1 + x2*w2 + x1*w3 + x3*w1          (dropped terms <= 2^-24 relative)
// in fp32 TMEM accumulators; bf16 x bf16 products are exact in fp32.  Six bf16 MMAs with K=16 cost the same
// number of tcgen05.mma instructions as three TF32 MMAs with K=8.
// TWO accumulators per tile: tensor-core accumulation truncates (about half an ulp of the running sum, biased, per
// tcgen05.mma), and with all six products in one accumulator that bias grows with 6 x (taps x K/16) instructions --
// measured 1e-6..1.5e-5 per conv in round 1, which pushed the encoder prediction to 1.35e-4.  Here the leading product
// x1*w1 accumulates in D_main, the five correction products (<= 2^-8 of it) in D_corr, whose truncation is 2^-8 smaller
// in absolute terms; the epilogue adds them with one rounded fp32 add.
//
// Same implicit GEMM as conv_tc.cu (M tile = TH x TW pixel patch, taps x 32-channel chunks on K, N tile <= 256),
// different operand plumbing:
//   * activations stay plain fp32 in HBM; four PRODUCER warps gather 16-byte pieces (bounds / padding by predicate,
//     stride-2 by address), split them and write three bf16 tiles into shared memory in the 64-byte-swizzled K-major
//     layout the tensor core reads (fence.proxy.async before the mbarrier arrive);
//   * weights are split once per step by the pack kernel into three bf16 planes [3][tap][N][K], one TMA per stage;
//   * one thread issues the products of a 16-channel K step -- products that share an A plane as ONE instruction over the
//     stacked weight planes (issue_kstep: 3 or 4 tcgen05.mma.kind::f16 instead of 6); epilogue warps read TMEM, add the
//     accumulator blocks and store fp32 (+bias, optional accumulate for the data gradient).
// Two kernels: conv_f32x3_kernel (any 1x1 / 3x3, stride 1 / 2: one gathered tile per tap) and conv_f32x3_halo_kernel
// (3x3 stride 1: ONE halo tile per patch and K chunk, taps as descriptor start rows, weights through their own ring).
// Forward and data gradient share them through a tap table, exactly like the bf16 kernels; the fp32 weight gradient runs on
// the bf16 wgrad kernels of conv_tc.cu through hi/lo planes (conv_wgrad_f32x2).
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace vae2 {
namespace t32 {

using namespace tc;

constexpr int kThreads = 320;       // warps 0-3 A producers | 4 MMA issuer | 5 weight TMA | 6-9 epilogue
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 200 * 1024;
constexpr int KC = 32;              // channels per stage: 32 bf16 = one 64-byte swizzled row
constexpr int kABytes = 128 * 64;   // one A tile (128 pixel rows x 64 bytes)

struct Params {
    int B, MH, MW, OH, OW, sO, oh_off, ow_off, sA;
    int AH, AW, lda, Ck;
    int Cn, ldo;
    int taps;
    signed char tap_dh[9], tap_dw[9], tap_w[9];
    int kchunks;
    int NT, n_tiles, TW, TH, tiles_w, tiles_h, total_tiles, stages, tmem_cols, acc_stages, acc_blocks, accumulate;
    const float* a;
    const float* bias;
    float* out;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// x = b1 + b2 + b3 exactly (each residual is exactly representable in fp32)
__device__ __forceinline__ void split3(float x, __nv_bfloat16& b1, __nv_bfloat16& b2, __nv_bfloat16& b3) {
    b1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b2);
    b3 = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}

// The six products of one 16-channel K step.  The issuing thread is the bottleneck of these kernels (in-step profile:
// ~90 cycles per tcgen05.mma whatever its N), so products that share an A plane are issued as ONE instruction whose B
// operand is the stack of weight planes (they lie back to back in shared memory, N = 2*NT or 3*NT rows):
//   blocks == 3, 3*NT <= 256 :  a1 x [w1 w2 w3] -> C0 C1 C2 ;  a2 x [w1 w2] -> C1 C2 ;  a3 x w1 -> C2          (3 instructions)
//   blocks == 3, 2*NT <= 256 :  a1 x [w1 w2] -> C0 C1 ;  a1 x w3 -> C2 ;  a2 x [w1 w2] -> C1 C2 ;  a3 x w1 -> C2  (4)
//   blocks == 2              :  the six products one by one, corrections (smallest first) -> C1, a1 x w1 -> C0      (6)
// C0 only ever receives the leading product a1*w1; the correction blocks are added to it in the epilogue.
__device__ __forceinline__ void issue_kstep(int blocks, uint32_t NT, uint32_t d0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t hi_a,
                                            uint32_t w1, uint32_t w_step, uint32_t hi_w, uint32_t idesc0, uint32_t first) {
    const uint32_t acc = first ? 0u : 1u;
    const uint32_t w2 = w1 + w_step, w3 = w2 + w_step;
    const uint32_t idesc1 = idesc0 | ((NT >> 3) << 17), idesc2 = idesc0 | ((2 * NT >> 3) << 17), idesc3 = idesc0 | ((3 * NT >> 3) << 17);
    if (blocks == 3 && 3 * NT <= 256) {
        umma_bf16_lohi(d0, a1, hi_a, w1, hi_w, idesc3, acc);
        umma_bf16_lohi(d0 + NT, a2, hi_a, w1, hi_w, idesc2, 1u);
        umma_bf16_lohi(d0 + 2 * NT, a3, hi_a, w1, hi_w, idesc1, 1u);
    } else if (blocks == 3) {
        umma_bf16_lohi(d0, a1, hi_a, w1, hi_w, idesc2, acc);
        umma_bf16_lohi(d0 + 2 * NT, a1, hi_a, w3, hi_w, idesc1, acc);
        umma_bf16_lohi(d0 + NT, a2, hi_a, w1, hi_w, idesc2, 1u);
        umma_bf16_lohi(d0 + 2 * NT, a3, hi_a, w1, hi_w, idesc1, 1u);
    } else {
        const uint32_t dc = d0 + NT;
        umma_bf16_lohi(dc, a3, hi_a, w1, hi_w, idesc1, acc);
        umma_bf16_lohi(dc, a1, hi_a, w3, hi_w, idesc1, 1u);
        umma_bf16_lohi(dc, a2, hi_a, w2, hi_w, idesc1, 1u);
        umma_bf16_lohi(dc, a2, hi_a, w1, hi_w, idesc1, 1u);
        umma_bf16_lohi(dc, a1, hi_a, w2, hi_w, idesc1, 1u);
        umma_bf16_lohi(d0, a1, hi_a, w1, hi_w, idesc1, acc);
    }
}

// epilogue warps 6..9 (TMEM -> registers -> fp32 global): D_main + D_corr (+bias) (+previous value)
__device__ __forceinline__ void epilogue_loop(const Params& p, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty,
                                          int warp, int lane) {
    const int per_img = p.tiles_w * p.tiles_h;
    const int total = p.total_tiles, gstride = gridDim.x, n_tiles = p.n_tiles;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ri = row / p.TW, rj = row % p.TW;
    const int two_acc = p.acc_stages == 2, NT = p.NT, Cn = p.Cn, accumulate = p.accumulate, blocks = p.acc_blocks;
    const float* bias = p.bias;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gstride, ++it) {
        const int as = two_acc ? (it & 1) : 0;
        const uint32_t use = two_acc ? (uint32_t)(it >> 1) : (uint32_t)it;
        const int nt = tile % n_tiles;
        const int pt = tile / n_tiles;
        const int b = pt / per_img;
        const int r = pt - b * per_img;
        const int gi = (r / p.tiles_w) * p.TH + ri, gj = (r % p.tiles_w) * p.TW + rj;
        const bool in_img = gi < p.MH && gj < p.MW;
        const int h = gi * p.sO + p.oh_off, w = gj * p.sO + p.ow_off;
        mbar_wait(&acc_full[as], use & 1);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * blocks * NT);
        float* orow = p.out + (((long long)b * p.OH + h) * p.OW + w) * p.ldo;
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t v[16], u[16];
            tmem_ld16(t0 + c0, v);               // leading product
            tmem_ld16(t0 + NT + c0, u);          // corrections
            if (blocks == 3) {
                uint32_t u2[16];
                tmem_ld16(t0 + 2 * NT + c0, u2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) + __uint_as_float(u2[i]));
            } else {
                tmem_ld_wait();
            }
            const int n = nt * NT + c0;
            if (in_img) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int nn = n + 4 * i;
                    if (nn < Cn) {                      // Cn is a multiple of 4
                        float4 f = make_float4(__uint_as_float(v[4 * i]) + __uint_as_float(u[4 * i]),
                                               __uint_as_float(v[4 * i + 1]) + __uint_as_float(u[4 * i + 1]),
                                               __uint_as_float(v[4 * i + 2]) + __uint_as_float(u[4 * i + 2]),
                                               __uint_as_float(v[4 * i + 3]) + __uint_as_float(u[4 * i + 3]));
                        if (bias != nullptr) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + nn));
                            f.x += bb.x; f.y += bb.y; f.z += bb.z; f.w += bb.w;
                        }
                        float4* dst = reinterpret_cast<float4*>(orow + nn);
                        if (accumulate) { const float4 o = *dst; f.x += o.x; f.y += o.y; f.z += o.z; f.w += o.w; }
                        *dst = f;
                    }
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_f32x3_kernel(const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t b_bytes = ((uint32_t)(p.NT * 64) + 1023u) & ~1023u;   // one weight plane tile (NT rows x 64 B)
    const uint32_t w_tx = 3u * (uint32_t)(p.NT * 64);                    // bytes one weight TMA delivers
    const uint32_t stage_bytes = 3 * kABytes + ((3u * (uint32_t)(p.NT * 64) + 1023u) & ~1023u);   // [A1][A2][A3][W1 W2 W3]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                       // [stages]  4 producer warps + 1 expect_tx arrive
    uint64_t* empty = bars + kMaxStages;         // [stages]  tcgen05.commit
    uint64_t* acc_full = bars + 2 * kMaxStages;  // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 5); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int per_img = p.tiles_w * p.tiles_h;
    const int total = p.total_tiles, gstride = gridDim.x, n_tiles = p.n_tiles, nstage = p.stages;
    const int ntaps = p.taps, kchunks = p.kchunks;

    if (warp < 4) {
        // ================= A producers: gather + split + swizzled store =================
        // Work item = (tile row, 8-channel group): consecutive threads take consecutive 32-byte pieces of a pixel's 128
        // bytes, so a warp's 16-byte load covers four whole pixel rows (was: one pixel per thread, 32 lines per request).
        // The loads of item i+1 are in flight while the slot of item i drains (an item = (tile, tap, 32-channel chunk)).
        const int tid = threadIdx.x;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gstride) {
            const int pt = tile / n_tiles;
            const int b = pt / per_img;
            const int r = pt - b * per_img;
            const int gi0 = (r / p.tiles_w) * p.TH, gj0 = (r % p.tiles_w) * p.TW;
            const float* img = p.a + (long long)b * p.AH * p.AW * p.lda;
            for (int tap = 0; tap < ntaps; ++tap) {
                const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
                long long off[4];
                bool okr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = (i * 128 + tid) >> 2;
                    const int gi = gi0 + row / p.TW, gj = gj0 + row % p.TW;
                    const int ih = gi * p.sA + dh, iw = gj * p.sA + dw;
                    okr[i] = gi < p.MH && gj < p.MW && ih >= 0 && ih < p.AH && iw >= 0 && iw < p.AW;
                    off[i] = ((long long)(okr[i] ? ih : 0) * p.AW + (okr[i] ? iw : 0)) * p.lda;
                }
                for (int kc = 0; kc < kchunks; ++kc) {
                    float4 v[4][2];
                    const int c = kc * KC + (tid & 3) * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float* src = img + off[i] + c;
                        v[i][0] = (okr[i] && c < p.Ck) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[i][1] = (okr[i] && c + 4 < p.Ck) ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* a1 = smem + (size_t)stage * stage_bytes;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = (i * 128 + tid) >> 2, j = tid & 3;
                        const float xs[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
                        uint4 o1, o2, o3;
                        __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(&o1);
                        __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(&o2);
                        __nv_bfloat16* h3 = reinterpret_cast<__nv_bfloat16*>(&o3);
#pragma unroll
                        for (int e = 0; e < 8; ++e) split3(xs[e], h1[e], h2[e], h3[e]);
                        const uint32_t o = (uint32_t)row * 64 + (((uint32_t)j ^ (uint32_t)((row >> 1) & 3)) << 4);   // 64-byte swizzle
                        *reinterpret_cast<uint4*>(a1 + o) = o1;
                        *reinterpret_cast<uint4*>(a1 + kABytes + o) = o2;
                        *reinterpret_cast<uint4*>(a1 + 2 * kABytes + o) = o3;
                    }
                    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[stage]);
                    if (++stage == nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ================= weight TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gstride) {
                const int nt = tile % n_tiles;
                for (int tap = 0; tap < ntaps; ++tap) {
                    const int wt = p.tap_w[tap];
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sb = smem + (size_t)stage * stage_bytes + 3 * kABytes;
                        mbar_expect_tx(&full[stage], w_tx);
                        tma_load_4d(sb, &map_w, &full[stage], kc * KC, nt * p.NT, wt, 0);     // box [32 ch][NT][1 tap][3 planes]
                        if (++stage == nstage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D = f32, A = B = bf16, K-major both, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);      // N is OR-ed in per instruction
            const uint32_t hi = desc_hi_word(64);
            const uint32_t lo_flags = 1u << 16;
            const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | lo_flags;
            const uint32_t stage_step = stage_bytes >> 4, a_step = kABytes >> 4, w_off = (3 * kABytes) >> 4;
            const uint32_t w_step = (uint32_t)(p.NT * 64) >> 4;          // planes are packed back to back by the TMA box
            const int nks = ntaps * kchunks, two_acc = p.acc_stages == 2, NT = p.NT, blocks = p.acc_blocks;
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gstride, ++it) {
                const int as = two_acc ? (it & 1) : 0;
                const uint32_t use = two_acc ? (uint32_t)(it >> 1) : (uint32_t)it;
                mbar_wait(&acc_empty[as], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(as * blocks * NT);
                uint32_t first = 1;
                for (int ks = 0; ks < nks; ++ks) {
                    // K steps of 16 channels in this chunk: the last chunk of a 36/72/144/272-lane tensor holds <= 16
                    const bool two_k = p.Ck - (ks % kchunks) * KC > 16;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a1 = a_lo0 + (uint32_t)stage * stage_step, a2 = a1 + a_step, a3 = a2 + a_step;
                    const uint32_t w1 = a1 + w_off;
                    issue_kstep(blocks, (uint32_t)NT, d0, a1, a2, a3, hi, w1, w_step, hi, idesc0, first);
                    if (two_k) issue_kstep(blocks, (uint32_t)NT, d0, a1 + 2, a2 + 2, a3 + 2, hi, w1 + 2, w_step, hi, idesc0, 0u);
                    first = 0;
                    umma_commit(&empty[stage]);
                    if (++stage == nstage) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        epilogue_loop(p, tmem_base, acc_full, acc_empty, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- 3x3 stride-1 layers: ONE halo tile per patch and K chunk -------------------------------------------------------------
// conv_f32x3_kernel gathers and splits the shifted tile of every tap: nine times the loads and nine times the split
// arithmetic for a 3x3 layer, and its four producer warps are what the kernel waits for (in-step profile r2, 64->64 3x3:
// 209 cycles per tcgen05.mma issued, against ~32 for the instruction itself).  Here the patch is 16 rows x 8 pixels and the
// producers split its 18 x 10 halo box ONCE per 32-channel chunk into three bf16 planes; a tap is then only a different
// start row of the K-major A descriptor (8-pixel row groups one box row apart, SBO = 10 rows; the swizzle is a function
// of the absolute shared-memory address -- same trick as conv_tc_halo_kernel, tools/probes/umma_shift_probe.cu).
// Weights stream through their own ring, one (chunk, tap) box of three planes per slot.
constexpr int kHTW = 8, kHTH = 16, kHBW = kHTW + 2, kHBH = kHTH + 2;
constexpr int kHaloRows = kHBW * kHBH;                       // 180 pixels
constexpr int kHaloPlane = (kHaloRows * 64 + 1023) / 1024 * 1024;
constexpr int kHaloItems = kHaloRows * 4;                    // (pixel, 8-channel group) work items of one stage
constexpr int kHaloIters = (kHaloItems + 127) / 128;
constexpr int kMaxWStages = 8, kMaxAStages = 3;

__global__ void __launch_bounds__(kThreads, 1)
conv_f32x3_halo_kernel(const __grid_constant__ CUtensorMap map_w, const Params p, const int a_stages, const int w_stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t w_tx = 3u * (uint32_t)(p.NT * 64);
    const uint32_t w_stage = (w_tx + 1023u) & ~1023u;
    const uint32_t a_stage = 3 * kHaloPlane;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem + (size_t)a_stages * a_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)w_stages * w_stage);
    uint64_t* a_full = bars;                            // [kMaxAStages]  4 producer warps
    uint64_t* a_empty = a_full + kMaxAStages;           // [kMaxAStages]  tcgen05.commit
    uint64_t* w_full = a_empty + kMaxAStages;           // [kMaxWStages]  expect_tx
    uint64_t* w_empty = w_full + kMaxWStages;           // [kMaxWStages]  tcgen05.commit
    uint64_t* acc_full = w_empty + kMaxWStages;         // [2]
    uint64_t* acc_empty = acc_full + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a_stages; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < w_stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int per_img = p.tiles_w * p.tiles_h;
    const int total = p.total_tiles, gstride = gridDim.x, kchunks = p.kchunks;

    if (warp < 4) {
        // ================= A producers: halo gather + split + swizzled store =================
        const int tid = threadIdx.x;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gstride) {
            const int b = tile / per_img;
            const int r = tile - b * per_img;
            const int h0 = (r / p.tiles_w) * kHTH - 1, w0 = (r % p.tiles_w) * kHTW - 1;
            const float* img = p.a + (long long)b * p.AH * p.AW * p.lda;
            for (int kc = 0; kc < kchunks; ++kc) {
                float4 v[kHaloIters][2];
#pragma unroll
                for (int i = 0; i < kHaloIters; ++i) {             // loads first: in flight while the slot drains
                    const int item = i * 128 + tid;
                    const int row = item >> 2, c = kc * KC + (item & 3) * 8;
                    const int ih = h0 + row / kHBW, iw = w0 + row % kHBW;
                    const bool ok = item < kHaloItems && ih >= 0 && ih < p.AH && iw >= 0 && iw < p.AW;
                    const float* src = img + ((long long)(ok ? ih : 0) * p.AW + (ok ? iw : 0)) * p.lda + c;
                    v[i][0] = (ok && c < p.Ck) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[i][1] = (ok && c + 4 < p.Ck) ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&a_empty[stage], phase ^ 1);
                uint8_t* a1 = smem + (size_t)stage * a_stage;
#pragma unroll
                for (int i = 0; i < kHaloIters; ++i) {
                    const int item = i * 128 + tid;
                    if (item < kHaloItems) {
                        const int row = item >> 2, j = item & 3;
                        const float xs[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
                        uint4 o1, o2, o3;
                        __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(&o1);
                        __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(&o2);
                        __nv_bfloat16* h3 = reinterpret_cast<__nv_bfloat16*>(&o3);
#pragma unroll
                        for (int e = 0; e < 8; ++e) split3(xs[e], h1[e], h2[e], h3[e]);
                        const uint32_t off = (uint32_t)row * 64 + (((uint32_t)j ^ (uint32_t)((row >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(a1 + off) = o1;
                        *reinterpret_cast<uint4*>(a1 + kHaloPlane + off) = o2;
                        *reinterpret_cast<uint4*>(a1 + 2 * kHaloPlane + off) = o3;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
                if (++stage == a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ================= weight TMA producer: one (chunk, tap) box of three planes per slot =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gstride)
                for (int kc = 0; kc < kchunks; ++kc)
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        mbar_expect_tx(&w_full[stage], w_tx);
                        tma_load_4d(wsm + (size_t)stage * w_stage, &map_w, &w_full[stage], kc * KC, 0, p.tap_w[tap], 0);
                        if (++stage == w_stages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
            const uint32_t hi_w = desc_hi_word(64);
            const uint32_t hi_a = ((uint32_t)(kHBW * 64) >> 4) | (1u << 14) | (4u << 29);     // SBO = one box row
            const uint32_t lo_flags = 1u << 16;
            const uint32_t a_base = smem_u32(smem), w_base = smem_u32(wsm);
            const uint32_t w_step = (uint32_t)(p.NT * 64) >> 4, a_step = (uint32_t)kHaloPlane >> 4;
            const int two_acc = p.acc_stages == 2, NT = p.NT, blocks = p.acc_blocks;
            int sa = 0, sw = 0, it = 0;
            uint32_t pa = 0, pw = 0;
            for (int tile = blockIdx.x; tile < total; tile += gstride, ++it) {
                const int as = two_acc ? (it & 1) : 0;
                const uint32_t use = two_acc ? (uint32_t)(it >> 1) : (uint32_t)it;
                mbar_wait(&acc_empty[as], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(as * blocks * NT);
                uint32_t first = 1;
                for (int kc = 0; kc < kchunks; ++kc) {
                    const bool two_k = p.Ck - kc * KC > 16;            // K steps of 16 channels in this chunk
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint32_t a_stage_addr = a_base + (uint32_t)sa * a_stage;
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&w_full[sw], pw);
                        tc_fence_after();
                        const uint32_t ta = a_stage_addr + (uint32_t)((p.tap_dh[tap] + 1) * kHBW + (p.tap_dw[tap] + 1)) * 64u;
                        const uint32_t a1 = ((ta >> 4) & 0x3FFFu) | lo_flags, a2 = a1 + a_step, a3 = a2 + a_step;
                        const uint32_t w1 = (((w_base + (uint32_t)sw * w_stage) >> 4) & 0x3FFFu) | lo_flags;
                        issue_kstep(blocks, (uint32_t)NT, d0, a1, a2, a3, hi_a, w1, w_step, hi_w, idesc0, first);
                        if (two_k) issue_kstep(blocks, (uint32_t)NT, d0, a1 + 2, a2 + 2, a3 + 2, hi_a, w1 + 2, w_step, hi_w, idesc0, 0u);
                        first = 0;
                        umma_commit(&w_empty[sw]);
                        if (++sw == w_stages) { sw = 0; pw ^= 1; }
                    }
                    umma_commit(&a_empty[sa]);
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        epilogue_loop(p, tmem_base, acc_full, acc_empty, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- weight split/pack: OIHW fp32 -> three bf16 planes [3][tap][Nf][Kf] (K contiguous, zero padded) -----------
__global__ void pack_weights_f32x3_kernel(const Tf32PackDesc* __restrict__ descs) {
    const Tf32PackDesc d = descs[blockIdx.y];
    const int kk = d.k * d.k;
    const int total = d.Cout * d.Cin * kk;
    const long long plane = (long long)kk * d.Nf * d.Kf, planeT = (long long)kk * d.NfT * d.KfT;
    __nv_bfloat16* fwd = reinterpret_cast<__nv_bfloat16*>(d.fwd);
    __nv_bfloat16* bwd = reinterpret_cast<__nv_bfloat16*>(d.bwd);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % kk;
        const int ci = (i / kk) % d.Cin;
        const int co = i / (kk * d.Cin);
        __nv_bfloat16 w1, w2, w3;
        split3(d.w[i], w1, w2, w3);
        const int pc = d.cin_map ? d.cin_map[ci] : ci;
        if (fwd) {        // forward operand: N = output channel, K = physical input lane
            const long long o = ((long long)tap * d.Nf + co) * d.Kf + pc;
            fwd[o] = w1; fwd[plane + o] = w2; fwd[2 * plane + o] = w3;
        }
        if (bwd) {        // data-gradient operand: N = physical input lane, K = output channel
            const long long o = ((long long)tap * d.NfT + pc) * d.KfT + co;
            bwd[o] = w1; bwd[planeT + o] = w2; bwd[2 * planeT + o] = w3;
        }
    }
}

struct Launch {
    const float* a; int lda, Ck, AH, AW;
    const void* w; int Nf, Kf, wtaps;           // three bf16 planes [3][wtaps][Nf][Kf]
    float* out; int ldo, OH, OW, Cn;
    const float* bias;
    int B, MH, MW, sO, oh_off, ow_off, sA;
    int ntaps; signed char dh[9], dw[9], wt[9];
    int accumulate;
};

}  // namespace t32

// N tiling shared by the packer (buffer sizes) and the launcher
static void t32_ntile(int Cn, int* n_tiles, int* NT) {
    const int npad = (Cn + 15) / 16 * 16;
    *n_tiles = (npad + 255) / 256;
    *NT = (((npad + *n_tiles - 1) / *n_tiles) + 15) / 16 * 16;
}

void conv_tf32_dims(const ConvGeom& g, int* Nf, int* Kf, int* NfT, int* KfT) {
    int nt, NT;
    t32_ntile(g.Cout_p, &nt, &NT);
    *Nf = nt * NT; *Kf = (g.Cin_p + 31) / 32 * 32;
    t32_ntile(g.Cin_p, &nt, &NT);
    *NfT = nt * NT; *KfT = (g.Cout_p + 31) / 32 * 32;
}

int conv_tf32_supported(const ConvGeom& g) {
    if (!(g.stride == 1 || (g.stride == 2 && g.k == 3))) return 0;
    if (!(g.k == 1 || g.k == 3)) return 0;
    if (g.Cin_p % 4 || g.Cout_p % 4 || g.ldx % 4 || g.ldy % 4) return 0;
    if (g.stride == 1 && (g.H != g.Ho || g.W != g.Wo)) return 0;
    if (g.stride == 2 && (g.Ho != (g.H + 1) / 2 || g.Wo != (g.W + 1) / 2)) return 0;
    return 1;
}

int pack_weights_tf32(const Tf32PackDesc* descs_dev, int n, cudaStream_t st) {
    if (n <= 0) return VAE2_OK;
    dim3 grid(8, n);
    t32::pack_weights_f32x3_kernel<<<grid, 256, 0, st>>>(descs_dev);
    return check_launch();
}

typedef CUresult (*EncodeTiledFn32)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn32 enc32() {
    static EncodeTiledFn32 fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn32>(sym);
    }
    return fn;
}

static int launch_t32(const t32::Launch& L, cudaStream_t st) {
    using namespace t32;
    EncodeTiledFn32 enc = enc32();
    if (enc == nullptr) return VAE2_ERR_UNSUPPORTED;
    if (L.MH <= 0 || L.MW <= 0) return VAE2_OK;
    Params p;
    p.B = L.B; p.MH = L.MH; p.MW = L.MW; p.OH = L.OH; p.OW = L.OW;
    p.sO = L.sO; p.oh_off = L.oh_off; p.ow_off = L.ow_off; p.sA = L.sA;
    p.AH = L.AH; p.AW = L.AW; p.lda = L.lda; p.Ck = L.Ck;
    p.Cn = L.Cn; p.ldo = L.ldo;
    p.taps = L.ntaps;
    for (int i = 0; i < 9; ++i) { p.tap_dh[i] = L.dh[i]; p.tap_dw[i] = L.dw[i]; p.tap_w[i] = L.wt[i]; }
    p.kchunks = L.Kf / KC;
    t32_ntile(L.Cn, &p.n_tiles, &p.NT);
    if (p.n_tiles * p.NT != L.Nf) return VAE2_ERR_ARG;
    p.accumulate = L.accumulate;
    p.a = L.a; p.bias = L.bias; p.out = L.out;
    // accumulator blocks per tile: leading product + one or two correction blocks (issue_kstep)
    p.acc_blocks = (2 * p.NT <= 256 && 3 * p.NT <= 512) ? 3 : 2;
    if (const char* e = getenv("VAE2_F32X3_MERGE")) { if (atoi(e) == 0) p.acc_blocks = 2; }
    p.acc_stages = (2 * p.acc_blocks * p.NT <= 512) ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.acc_stages * p.acc_blocks * p.NT) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return VAE2_ERR_UNSUPPORTED;

    CUtensorMap map_w;
    {
        cuuint64_t dims[4] = {(cuuint64_t)L.Kf, (cuuint64_t)L.Nf, (cuuint64_t)L.wtaps, 3};
        cuuint64_t strides[3] = {(cuuint64_t)L.Kf * 2, (cuuint64_t)L.Nf * L.Kf * 2, (cuuint64_t)L.wtaps * L.Nf * L.Kf * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.NT, 1, 3};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(L.w), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return VAE2_ERR_ARG;
    }
    // ---- halo path: 3x3, unit strides, one N tile ----
    {
        bool halo = L.ntaps == 9 && L.sA == 1 && L.sO == 1 && L.oh_off == 0 && L.ow_off == 0 && p.n_tiles == 1 &&
                    L.MH == L.AH && L.MW == L.AW && L.MH >= 4 && L.MW >= 4;
        for (int i = 0; halo && i < 9; ++i) halo = L.dh[i] >= -1 && L.dh[i] <= 1 && L.dw[i] >= -1 && L.dw[i] <= 1;
        if (const char* e = getenv("VAE2_F32X3_HALO")) { if (atoi(e) == 0) halo = false; }
        const int w_stage = (3 * p.NT * 64 + 1023) / 1024 * 1024;
        const int bar_bytes = 1024 + (2 * kMaxAStages + 2 * kMaxWStages + 4) * 8 + 16;
        const int budget = 227 * 1024 - bar_bytes;
        int a_stages = kMaxAStages;
        while (a_stages > 2 && (budget - a_stages * 3 * kHaloPlane) / w_stage < 4) --a_stages;
        int w_stages = (budget - a_stages * 3 * kHaloPlane) / w_stage;
        if (w_stages > kMaxWStages) w_stages = kMaxWStages;
        if (halo && w_stages >= 2) {
            p.TW = kHTW; p.TH = kHTH;
            p.tiles_w = (L.MW + kHTW - 1) / kHTW;
            p.tiles_h = (L.MH + kHTH - 1) / kHTH;
            p.total_tiles = L.B * p.tiles_w * p.tiles_h;
            p.stages = a_stages;
            const size_t smem = (size_t)a_stages * 3 * kHaloPlane + (size_t)w_stages * w_stage + bar_bytes;
            static bool attr_set_h = false;
            if (!attr_set_h) {
                if (cudaFuncSetAttribute(conv_f32x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                    return VAE2_ERR_CUDA;
                attr_set_h = true;
            }
            const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
            note_kernel("t32::conv_f32x3_halo_kernel");
            conv_f32x3_halo_kernel<<<grid, kThreads, smem, st>>>(map_w, p, a_stages, w_stages);
            return check_launch();
        }
    }
    int tw = 128;
    while (tw > 8 && tw / 2 >= L.MW) tw >>= 1;
    p.TW = tw; p.TH = 128 / tw;
    p.tiles_w = (L.MW + p.TW - 1) / p.TW;
    p.tiles_h = (L.MH + p.TH - 1) / p.TH;
    p.total_tiles = L.B * p.tiles_w * p.tiles_h * p.n_tiles;
    const int stage_bytes = 3 * kABytes + (3 * p.NT * 64 + 1023) / 1024 * 1024;
    int stages = kSmemBudget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return VAE2_ERR_UNSUPPORTED;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * kMaxStages + 4) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_f32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return VAE2_ERR_CUDA;
        attr_set = true;
    }
    int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    note_kernel("t32::conv_f32x3_kernel");
    conv_f32x3_kernel<<<grid, kThreads, smem, st>>>(map_w, p);
    return check_launch();
}

// w_fwd: bf16 planes [3][tap][Nf][Kf] as produced by pack_weights_tf32 (fwd operand)
int conv_fwd_tf32(const void* x, const void* w_fwd, const float* bias, void* y, const ConvGeom& g, cudaStream_t st) {
    if (!conv_tf32_supported(g)) return VAE2_ERR_UNSUPPORTED;
    int Nf, Kf, NfT, KfT;
    conv_tf32_dims(g, &Nf, &Kf, &NfT, &KfT);
    t32::Launch L{};
    L.a = (const float*)x; L.lda = g.ldx; L.Ck = g.Cin_p; L.AH = g.H; L.AW = g.W;
    L.w = w_fwd; L.Nf = Nf; L.Kf = Kf; L.wtaps = g.k * g.k;
    L.out = (float*)y; L.ldo = g.ldy; L.OH = g.Ho; L.OW = g.Wo; L.Cn = g.Cout_p; L.bias = bias;
    L.B = g.B; L.MH = g.Ho; L.MW = g.Wo; L.sO = 1; L.oh_off = 0; L.ow_off = 0; L.sA = g.stride;
    L.ntaps = g.k * g.k;
    for (int t = 0; t < L.ntaps; ++t) { L.dh[t] = (signed char)(t / g.k - g.pad); L.dw[t] = (signed char)(t % g.k - g.pad); L.wt[t] = (signed char)t; }
    L.accumulate = 0;
    return launch_t32(L, st);
}

// w_bwd: bf16 planes [3][tap][NfT][KfT] (data-gradient operand).  Stride 2 -> four output parity classes (see conv_tc.cu).
int conv_dgrad_tf32(const void* dy, const void* w_bwd, void* dx, const ConvGeom& g, int accumulate, cudaStream_t st) {
    if (!conv_tf32_supported(g)) return VAE2_ERR_UNSUPPORTED;
    int Nf, Kf, NfT, KfT;
    conv_tf32_dims(g, &Nf, &Kf, &NfT, &KfT);
    t32::Launch L{};
    L.a = (const float*)dy; L.lda = g.ldy; L.Ck = g.Cout_p; L.AH = g.Ho; L.AW = g.Wo;
    L.w = w_bwd; L.Nf = NfT; L.Kf = KfT; L.wtaps = g.k * g.k;
    L.out = (float*)dx; L.ldo = g.ldx; L.OH = g.H; L.OW = g.W; L.Cn = g.Cin_p; L.bias = nullptr;
    L.B = g.B; L.sA = 1; L.accumulate = accumulate;
    if (g.stride == 1) {
        L.MH = g.H; L.MW = g.W; L.sO = 1; L.oh_off = 0; L.ow_off = 0;
        L.ntaps = g.k * g.k;
        for (int t = 0; t < L.ntaps; ++t) { L.dh[t] = (signed char)(g.pad - t / g.k); L.dw[t] = (signed char)(g.pad - t % g.k); L.wt[t] = (signed char)t; }
        return launch_t32(L, st);
    }
    for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
            L.MH = (g.H - ph + 1) / 2; L.MW = (g.W - pw + 1) / 2;
            L.sO = 2; L.oh_off = ph; L.ow_off = pw;
            int n = 0;
            for (int ky = 0; ky < 3; ++ky) {
                if (((ph + 1 - ky) & 1) != 0) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    if (((pw + 1 - kx) & 1) != 0) continue;
                    L.dh[n] = (signed char)((ph + 1 - ky) / 2); L.dw[n] = (signed char)((pw + 1 - kx) / 2);
                    L.wt[n] = (signed char)(ky * 3 + kx);
                    ++n;
                }
            }
            L.ntaps = n;
            if (int e = launch_t32(L, st)) return e;
        }
    return VAE2_OK;
}

}  // namespace vae2
