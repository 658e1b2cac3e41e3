// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// Replaces cuDNN behind the reference's stride-1 nn.Conv2d sites (3x3 p1 and 1x1;
// lib/models/enc_hrnet.py:27-30, 38-41, 70-76, 188-195, 324-337, 381-390, 1010-1017) on the bf16
// path, for the forward pass and -- with the transposed weight pack and mirrored tap offsets -- for
// the data gradient.  (Stride-2 convs, 8 % of the MACs, and the weight gradient use conv_simt.cu.)
//
//   D[pixel][n] = sum_{tap} sum_{c}  A[pixel + off(tap)][c] * Wq[tap][n][c]
//
// Mapping onto the hardware:
//   * M tile = 128 output pixels = a TH x TW patch of one image (TW*TH = 128), which is one TMA
//     box [KC ch][TW][TH][1] of the channels-last activation shifted by the tap offset; TMA's
//     out-of-bounds zero fill IS the conv padding.  The box lands in shared memory as 128 rows of
//     KC*2 bytes with the 128/64/32-byte swizzle, i.e. exactly the K-major canonical layout a
//     tcgen05 shared-memory descriptor reads.
//   * N tile = NT output channels (<= 256): a TMA box [KC][NT][1] of the packed weights.
//   * one elected thread issues tcgen05.mma (M=128, N=NT, K=16) per 16 channels; the fp32
//     accumulator (128 lanes x NT columns) lives in TMEM, double buffered so the epilogue of tile
//     i overlaps the MMAs of tile i+1.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
//     (tcgen05.ld 32 lanes x 16 columns -> +bias -> bf16 -> 32-byte global stores, one pixel row
//     per thread).  Persistent CTAs (one per SM) walk the tile list.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace vae2 {

namespace tc {

constexpr int kThreads = 192;          // 6 warps
constexpr int kEpiWarp0 = 2;           // first epilogue warp
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;

struct Params {
    int B;
    int MH, MW;               // extent of the M grid (one GEMM row per grid point (i, j))
    int OH, OW;               // spatial extent of the output tensor
    int sO, oh_off, ow_off;   // output pixel of grid point (i, j) = (i*sO + oh_off, j*sO + ow_off)
    int sA;                   // A box origin = (j0*sA + dw, i0*sA + dh); 2 for stride-2 forward (map has element stride 2)
    int Cn;                   // output channel lanes (N extent), multiple of 16
    int ldo;                  // output pixel pitch (elements)
    int taps;                 // number of (dh, dw, weight-tap) entries
    signed char tap_dh[9], tap_dw[9], tap_w[9];
    int kchunks;              // K chunks of KC channels per tap
    int KC;                   // 16 / 32 / 64 channels per stage
    int NT;                   // N tile (multiple of 16, <= 256)
    int n_tiles;              // tiles along N
    int TW, TH;               // pixel patch, TW*TH = 128
    int tiles_w, tiles_h;     // patches per image
    int total_tiles;
    int stages;
    int tmem_cols;            // power of two >= 2*NT (or >= NT when single buffered)
    int acc_stages;           // 1 or 2
    int accumulate;           // epilogue adds to the existing output (dgrad +=)
    const float* bias;        // per output channel (padded to Cn) or null
    __nv_bfloat16* out;
};

__global__ void __launch_bounds__(kThreads, 4)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][A tile | B tile] (1024-aligned), then barriers
    const uint32_t row_bytes = p.KC * 2;
    const uint32_t a_bytes = 128 * row_bytes;
    const uint32_t b_bytes = ((p.NT * row_bytes + 1023) / 1024) * 1024;
    const uint32_t stage_bytes = a_bytes + b_bytes;   // a_bytes is a multiple of 1024 for KC >= 16? 128*32 = 4096 yes
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                       // [stages]
    uint64_t* empty = bars + kMaxStages;         // [stages]
    uint64_t* acc_full = bars + 2 * kMaxStages;  // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int pt = tile / p.n_tiles;
                const int b = pt / per_img;
                const int r = pt - b * per_img;
                const int h0 = (r / p.tiles_w) * p.TH, w0 = (r % p.tiles_w) * p.TW;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int dh = p.tap_dh[tap], dw = p.tap_dw[tap], wt = p.tap_w[tap];
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full[stage], a_bytes + p.NT * row_bytes);
                        tma_load_4d(sa, &map_a, &full[stage], kc * p.KC, w0 * p.sA + dw, h0 * p.sA + dh, b);
                        tma_load_3d(sa + a_bytes, &map_w, &full[stage], kc * p.KC, nt * p.NT, wt);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int as = p.acc_stages == 2 ? (it & 1) : 0;
                const uint32_t use = p.acc_stages == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
                mbar_wait(&acc_empty[as], (use & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.NT);
                uint32_t accum = 0;
                for (int ks = 0; ks < p.taps * p.kchunks; ++ks) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
                    for (int k = 0; k < p.KC / 16; ++k) {
                        umma_bf16(d_tmem, make_desc(sa + k * 32, row_bytes), make_desc(sb + k * 32, row_bytes), idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&empty[stage]);                 // frees the smem stage when these MMAs finish
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);                     // accumulator complete -> epilogue
            }
        }
    } else {
        // ================= epilogue warps (TMEM -> registers -> global) =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;          // tile row == TMEM lane == pixel within the patch
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int as = p.acc_stages == 2 ? (it & 1) : 0;
            const uint32_t use = p.acc_stages == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
            const int nt = tile % p.n_tiles;
            const int pt = tile / p.n_tiles;
            const int b = pt / per_img;
            const int r = pt - b * per_img;
            const int gi = (r / p.tiles_w) * p.TH + row / p.TW, gj = (r % p.tiles_w) * p.TW + row % p.TW;
            const bool in_img = gi < p.MH && gj < p.MW;
            const int h = gi * p.sO + p.oh_off, w = gj * p.sO + p.ow_off;
            mbar_wait(&acc_full[as], use & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.NT);
            __nv_bfloat16* orow = p.out + (((long long)b * p.OH + h) * p.OW + w) * p.ldo;
            for (int c0 = 0; c0 < p.NT; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t0 + c0, v);
                tmem_ld_wait();
                const int n = nt * p.NT + c0;
                if (in_img && n < p.Cn) {
                    const bool half = n + 8 >= p.Cn;          // 8-lane tail (prediction heads): only one 16-byte store
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] += (i < 8 || !half) ? __ldg(p.bias + n + i) : 0.f;
                    }
                    uint4* dst = reinterpret_cast<uint4*>(orow + n);
                    if (p.accumulate) {
                        uint4 o[2];
                        o[0] = dst[0];
                        o[1] = half ? make_uint4(0u, 0u, 0u, 0u) : dst[1];
                        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(o);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { float2 t = __bfloat1622float2(oh[i]); f[2 * i] += t.x; f[2 * i + 1] += t.y; }
                    }
                    uint4 o[2];
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
                    for (int i = 0; i < 8; ++i) oh[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                    dst[0] = o[0];
                    if (!half) dst[1] = o[1];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);     // 4 epilogue warps -> count 4
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- 3x3 stride-1 with ONE halo tile per (pixel patch, K chunk) ---------------------------------------------
// The nine taps of a 3x3 convolution read nine shifted views of the same pixels.  conv_tc_kernel fetches each
// view with its own TMA (9x the activation bytes through L2 -> shared memory, which is what bounds it for
// narrow layers); here the patch is 16 rows x 8 pixels and ONE TMA box [KC][10][18] brings the patch plus its
// 1-pixel halo.  Tap (dh, dw) is then only a different START ADDRESS of the A descriptor: 8-row groups are the
// 8 pixels of one patch row (contiguous in the box), consecutive groups are one box row = 10 pixels apart (SBO),
// and the swizzle is a function of the absolute shared-memory address, so a start that is not aligned to the
// swizzle atom reads back exactly what TMA wrote (tools/probes/umma_shift_probe.cu checks this on the device).
// The packed weights of all taps stay resident in shared memory for the CTA's lifetime; a pipeline stage is a
// halo tile, so a patch costs kchunks barrier round trips instead of 9*kchunks.
constexpr int kHaloTW = 8, kHaloTH = 16, kHaloBW = kHaloTW + 2, kHaloBH = kHaloTH + 2;

__global__ void __launch_bounds__(kThreads, 3)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t row_bytes = p.KC * 2;
    const uint32_t a_bytes = kHaloBW * kHaloBH * row_bytes;                 // TMA transaction size of one halo tile
    const uint32_t a_stage = (a_bytes + 1023) / 1024 * 1024;
    const uint32_t atom = 8 * row_bytes;                                    // swizzle repeat
    const uint32_t w_block = (p.NT * row_bytes + atom - 1) / atom * atom;  // one (K chunk, tap) weight tile
    const uint32_t w_bytes = ((uint32_t)(p.kchunks * p.taps) * w_block + 1023) / 1024 * 1024;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;
    uint8_t* asm_ = smem + w_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(asm_ + (size_t)p.stages * a_stage);
    uint64_t* full = bars;                       // [stages]
    uint64_t* empty = bars + kMaxStages;         // [stages]
    uint64_t* acc_full = bars + 2 * kMaxStages;  // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint64_t* w_full = acc_empty + 2;            // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(p.kchunks * p.taps) * (uint32_t)p.NT * row_bytes);
            for (int kc = 0; kc < p.kchunks; ++kc)
                for (int tap = 0; tap < p.taps; ++tap)
                    tma_load_3d(wsm + (size_t)(kc * p.taps + tap) * w_block, &map_w, w_full, kc * p.KC, 0, p.tap_w[tap]);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int b = tile / per_img;
                const int r = tile - b * per_img;
                const int h0 = (r / p.tiles_w) * kHaloTH, w0 = (r % p.tiles_w) * kHaloTW;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], a_bytes);
                    tma_load_4d(asm_ + (size_t)stage * a_stage, &map_a, &full[stage], kc * p.KC, w0 - 1, h0 - 1, b);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
            // descriptor high words: SBO = one box row (A) / 8 rows (weights), version 1, swizzle by span
            const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
            const uint32_t hi_a = ((kHaloBW * row_bytes) >> 4) | (1u << 14) | (layout << 29);
            const uint32_t hi_w = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
            mbar_wait(w_full, 0);
            tc_fence_after();
            const uint32_t w0s = smem_u32(wsm);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                mbar_wait(&acc_empty[as], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.NT);
                uint32_t accum = 0;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(asm_ + (size_t)stage * a_stage);
                    const uint32_t sw = w0s + (uint32_t)(kc * p.taps) * w_block;
                    for (int tap = 0; tap < p.taps; ++tap) {
                        const uint32_t ta = sa + (uint32_t)((p.tap_dh[tap] + 1) * kHaloBW + (p.tap_dw[tap] + 1)) * row_bytes;
                        const uint32_t tw = sw + (uint32_t)tap * w_block;
                        for (int k = 0; k < p.KC / 16; ++k) {
                            const uint64_t da = ((uint64_t)hi_a << 32) | (1u << 16) | (((ta + k * 32) >> 4) & 0x3FFF);
                            const uint64_t db = ((uint64_t)hi_w << 32) | (1u << 16) | (((tw + k * 32) >> 4) & 0x3FFF);
                            umma_bf16(d_tmem, da, db, idesc, accum);
                            accum = 1;
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        // ================= epilogue warps (TMEM -> registers -> global) =================
        const int q = warp & 3;
        const int row = q * 32 + lane;          // tile row == TMEM lane == pixel (row / 8, row % 8) of the patch
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            const int b = tile / per_img;
            const int r = tile - b * per_img;
            const int h = (r / p.tiles_w) * kHaloTH + row / kHaloTW, w = (r % p.tiles_w) * kHaloTW + row % kHaloTW;
            const bool in_img = h < p.MH && w < p.MW;
            mbar_wait(&acc_full[as], use & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.NT);
            __nv_bfloat16* orow = p.out + (((long long)b * p.OH + h) * p.OW + w) * p.ldo;
            for (int c0 = 0; c0 < p.NT; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t0 + c0, v);
                tmem_ld_wait();
                if (in_img && c0 < p.Cn) {
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + c0 + i);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(orow + c0);
                    if (p.accumulate) {
                        uint4 o[2] = {dst[0], dst[1]};
                        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(o);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { float2 t = __bfloat1622float2(oh[i]); f[2 * i] += t.x; f[2 * i + 1] += t.y; }
                    }
                    uint4 o[2];
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
                    for (int i = 0; i < 8; ++i) oh[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                    dst[0] = o[0];
                    dst[1] = o[1];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static CUtensorMapSwizzle swz(int kc) {
    return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// K chunk (channels per pipeline stage).  A stage costs the single issuing thread ~300 cycles of barrier
// wait / fence / commit plus ~110 cycles per tcgen05.mma (measured), so fewer, fatter stages win even when
// the last chunk is partly empty: TMA zero-fills the channels beyond the tensor's extent for free.
static int pick_kc(int c) {
    int best = 16, best_cost = 1 << 30;
    for (int kc = 16; kc <= 64; kc *= 2) {
        if (kc > 16 && c <= kc / 2) break;
        const int cost = ((c + kc - 1) / kc) * (300 + (kc / 16) * 110);
        if (cost < best_cost) { best_cost = cost; best = kc; }
    }
    return best;
}

static int next_pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}


// =================================================================================================
// Weight gradient on tcgen05:   dW[tap][ci][co] = sum_pixels  x[pixel + off(tap)][ci] * dy[pixel][co]
//
// GEMM with K = pixels.  Both operands are read exactly as they lie in memory (channels-last), i.e.
// MN-major: a TMA box [atom channels][TW][TH] lands as KP = TW*TH rows (pixels) of `atom` channels,
// which is the MN-major canonical layout (k rows of one swizzle span, 8-row groups SBO apart, the
// next `atom` channels LBO = KP*row_bytes further on).  A 5-D tensor map (c_lo, w, h, b, c_hi)
// fetches ALL channel atoms of one tap in a single TMA.
//   A (M side) : rows m = tap*Cin_p + ci  -> the 9 tap-shifted x tiles stacked along M; one
//                tcgen05.mma covers 128 consecutive rows ("group"), possibly spanning taps.
//   B (N side) : dy tile, N = Cout_p, loaded once per pixel chunk and shared by all taps.
//   D          : groups x [128 lanes x Cout_p columns] fp32 in TMEM (<= 512 columns per CTA; more
//                groups -> several "sets" handled by different CTAs).
// Split-K: CTA (set, range) walks pixel tiles range, range+nranges, ...; partial sums go to a
// workspace slot per range and a second kernel adds the slots into dwp[tap][Cin_p][Cout_p].
// =================================================================================================
struct WParams {
    int B, H, W;
    int Cin_p, Cout_p, taps, ksz, pad;
    int atomA, atomB;          // channels per swizzle atom on the M / N side (16, 32 or 64)
    int TW, TH, KP;            // pixel patch per stage (KP = TW*TH, multiple of 16)
    int tiles_w, tiles_h, total_ptiles;
    int M_total;               // taps * Cin_p
    int set_groups, nsets, nranges;
    int sA;                    // x box origin = (w0*sA + kx - pad, h0*sA + ky - pad); 2 for stride-2 convs
    int NT, n_tiles;           // N tile (<= 256 output channels) and tile count
    int a_atoms_stage;         // atoms reserved for A per stage
    int stages;
    int tmem_cols;
    int dual;                  // fp32 gradient from bf16 hi/lo planes in ONE pass: A = {x_hi, x_lo}, B = {dy_hi, dy_lo};
                               // per group two accumulators: x_hi*dy_hi | x_hi*dy_lo + x_lo*dy_hi, added in the epilogue
    float* ws;                 // [nranges][M_total][Cout_p]
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// MN-major descriptor: rows (k) of `row_bytes`, 8-row groups SBO = 8*row_bytes apart, atoms LBO apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                const __grid_constant__ CUtensorMap map_xl, const __grid_constant__ CUtensorMap map_dyl, const WParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t rowA = p.atomA * 2, rowB = p.atomB * 2;
    const uint32_t atomA_bytes = p.KP * rowA, atomB_bytes = p.KP * rowB;     // multiples of 1024 for KP = 32/64... (>= 512)
    const uint32_t a_bytes = (uint32_t)p.a_atoms_stage * atomA_bytes;
    const uint32_t b_bytes = (uint32_t)(p.NT / p.atomB) * atomB_bytes;
    const uint32_t planes = p.dual ? 2u : 1u;                                // stage = [A_hi][A_lo][B_hi][B_lo]
    const uint32_t stage_bytes = ((planes * (a_bytes + b_bytes) + 1023) / 1024) * 1024;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* acc_full = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int set = blockIdx.x % p.nsets;
    const int ntile = (blockIdx.x / p.nsets) % p.n_tiles;
    const int range = blockIdx.x / (p.nsets * p.n_tiles);
    const int n0 = ntile * p.NT;
    const int g_lo = set * p.set_groups;
    const int groups_total = (p.M_total + 127) / 128;
    const int g_hi = min(groups_total, g_lo + p.set_groups);      // exclusive
    const int m_lo = g_lo * 128, m_hi = min(p.M_total, g_hi * 128);
    const int t0 = m_lo / p.Cin_p, t1 = (m_hi - 1) / p.Cin_p;      // taps this set touches
    const int nchunkA = p.Cin_p / p.atomA;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_img = p.tiles_w * p.tiles_h;
    const int n_my_tiles = (p.total_ptiles - range + p.nranges - 1) / p.nranges;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = planes * ((uint32_t)(t1 - t0 + 1) * nchunkA * atomA_bytes + b_bytes);
            for (int i = 0; i < n_my_tiles; ++i) {
                const int pt = range + i * p.nranges;
                const int b = pt / per_img;
                const int r = pt - b * per_img;
                const int h0 = (r / p.tiles_w) * p.TH, w0 = (r % p.tiles_w) * p.TW;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], tx_bytes);
                for (int tap = t0; tap <= t1; ++tap) {
                    const int ky = tap / p.ksz, kx = tap - ky * p.ksz;
                    tma_load_5d(sa + (size_t)(tap - t0) * nchunkA * atomA_bytes, &map_x, &full[stage], 0,
                                w0 * p.sA + kx - p.pad, h0 * p.sA + ky - p.pad, b, 0);
                    if (p.dual)
                        tma_load_5d(sa + a_bytes + (size_t)(tap - t0) * nchunkA * atomA_bytes, &map_xl, &full[stage], 0,
                                    w0 * p.sA + kx - p.pad, h0 * p.sA + ky - p.pad, b, 0);
                }
                tma_load_5d(sa + planes * a_bytes, &map_dy, &full[stage], 0, w0, h0, b, n0 / p.atomB);
                if (p.dual) tma_load_5d(sa + 2 * a_bytes + b_bytes, &map_dyl, &full[stage], 0, w0, h0, b, n0 / p.atomB);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D=f32, A=B=bf16, A and B MN-major (bits 15, 16), N = Cout_p, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(2 * p.NT >> 3) << 17);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_my_tiles; ++i) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t sb = sa + planes * a_bytes;
                for (int g = g_lo; g < g_hi; ++g) {
                    const uint32_t a_rel = (uint32_t)(g * 128 - t0 * p.Cin_p) / (uint32_t)p.atomA;   // first atom of this group
                    const uint32_t d_tmem = tmem_base + (uint32_t)((g - g_lo) * p.NT) * planes;
                    for (int k = 0; k < p.KP / 16; ++k) {
                        const uint64_t da = make_desc_mn(sa + a_rel * atomA_bytes + k * 16 * rowA, rowA, atomA_bytes);
                        const uint64_t db = make_desc_mn(sb + k * 16 * rowB, rowB, atomB_bytes);
                        const uint32_t acc = (i | k) ? 1u : 0u;
                        if (!p.dual) {
                            umma_bf16(d_tmem, da, db, idesc, acc);
                            continue;
                        }
                        // x_hi * [dy_hi | dy_lo]: the two dy planes lie back to back, so one instruction of N = 2*NT
                        // fills both accumulators (two instructions when 2*NT > 256); then x_lo * dy_hi into the second
                        const uint64_t dal = make_desc_mn(sa + a_bytes + a_rel * atomA_bytes + k * 16 * rowA, rowA, atomA_bytes);
                        if (2 * p.NT <= 256) {
                            umma_bf16(d_tmem, da, db, idesc2, acc);
                        } else {
                            umma_bf16(d_tmem, da, db, idesc, acc);
                            umma_bf16(d_tmem + (uint32_t)p.NT, da, make_desc_mn(sb + b_bytes + k * 16 * rowB, rowB, atomB_bytes), idesc, acc);
                        }
                        umma_bf16(d_tmem + (uint32_t)p.NT, dal, db, idesc, 1u);
                    }
                }
                umma_commit(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float* ws = p.ws + (size_t)range * p.M_total * p.Cout_p;
        for (int g = g_lo; g < g_hi; ++g) {
            const int m = g * 128 + q * 32 + lane;
            const uint32_t t0a = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g_lo) * p.NT) * planes;
            for (int c0 = 0; c0 < p.NT; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t0a + c0, v);
                if (p.dual) {
                    uint32_t u[16];
                    tmem_ld16(t0a + p.NT + c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
                } else {
                    tmem_ld_wait();
                }
                if (m < p.M_total && n0 + c0 < p.Cout_p) {
                    float4* dst = reinterpret_cast<float4*>(ws + (size_t)m * p.Cout_p + n0 + c0);
                    const int nq = min(4, (p.Cout_p - n0 - c0) / 4);        // 8-lane outputs: two float4 of the 16-column group
                    if (n_my_tiles > 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i < nq)
                                dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                     __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i < nq) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// dwp[i] = sum over split-K slots.  COLS float4 columns x (256 / COLS) slot groups per CTA: a thread adds every
// (256/COLS)-th slot (independent loads, a short dependent chain), the groups are folded through shared memory in a fixed
// order.  Small gradients (18 -> 18: 9216 floats in 296 slots) take COLS = 8 so that 4x more CTAs share the slots -- with
// 32 columns only 72 CTAs ran 37 dependent rounds each (ncu r2f: 14.4 us for 11 MB).
template <int COLS>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dwp, long long n, int nslots) {
    constexpr int GROUPS = 256 / COLS;
    __shared__ float4 sm[GROUPS][COLS];
    const int ex = threadIdx.x % COLS, sg = threadIdx.x / COLS;
    const long long n4 = n / 4;
    const long long i = (long long)blockIdx.x * COLS + ex;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
        const float4* src = reinterpret_cast<const float4*>(ws) + i;
#pragma unroll 4
        for (int s = sg; s < nslots; s += GROUPS) {
            const float4 v = src[(long long)s * n4];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    }
    sm[sg][ex] = a;
    __syncthreads();
    if (sg == 0 && i < n4) {
#pragma unroll
        for (int g = 1; g < GROUPS; ++g) {
            const float4 v = sm[g][ex];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4*>(dwp)[i] = a;
    }
}

static void launch_wgrad_reduce(const float* ws, float* dwp, long long n, int nslots, cudaStream_t st) {
    const long long n4 = n / 4;
    if (n4 < 2LL * kNumSMs * 32 && nslots >= 64)
        wgrad_reduce_kernel<8><<<(unsigned)((n4 + 7) / 8), 256, 0, st>>>(ws, dwp, n, nslots);
    else
        wgrad_reduce_kernel<32><<<(unsigned)((n4 + 31) / 32), 256, 0, st>>>(ws, dwp, n, nslots);
}

// ---- weight gradient of 3x3 stride-1 layers with ONE halo tile per pixel patch --------------------------------------------
// wgrad_tc_kernel fetches the nine tap-shifted x tiles of a patch separately: 9x the activation bytes through L2 -> shared
// memory (ncu r2a, 18->18: 335 MB for a 67 MB problem; the kernel sits at 11 % of its HBM floor).  Here the patch is 16 rows
// x 8 pixels and ONE TMA box [Cin_p][10][18] brings it with its halo.  The operands stay MN-major (rows = pixels = K), and
// a tap is a descriptor, not a copy:
//   * kernel row ky   -> start address shifted by ky box rows;
//   * the two 8-pixel K groups of a 16-pixel K step are consecutive patch rows, one BOX row apart (SBO = 10 pixels);
//   * the taps kx = 0,1,2 of one kernel row are stacked on the M axis as consecutive "atoms" ONE PIXEL apart (LBO = one
//     row): M = 128 covers 4 atoms of 32 channels (the fourth, kx = 3, is junk and dropped in the epilogue) or 2 atoms of 64.
// tools/probes/umma_mn_shift_probe.cu established on the device that MN-major descriptors take unaligned starts, an SBO of a
// box row and overlapping atoms exactly (profiles/r2_probe_mn.txt).  Accumulators: one TMEM region per kernel row (and per
// atom pair for 64 channels); split-K over patches with per-range partials, folded by wgrad_reduce_kernel as before.
struct WHParams {
    int B, H, W;
    int Cin_p, Cout_p;          // Cin_p = 32, or a multiple of 64 handled as `csets` chunks of Cc = 64 lanes by different CTAs
    int Cc, csets;              // lanes per chunk (one swizzle atom per pixel row) and chunk count; Cout_p <= 256
    int atomB, NT;
    int tiles_w, tiles_h, total_ptiles, nranges;
    int mmas_per_row;           // 1 (Cin_p = 32: kx 0..3 in one M=128) or 2 (Cin_p = 64: kx {0,1} and {2,3})
    int stages, tmem_cols;
    int dual;                   // hi/lo planes of an fp32 gradient in one pass (see WParams::dual)
    float* ws;                  // [nranges][9][Cin_p][Cout_p]
};

__device__ __forceinline__ uint64_t make_desc_mn_ex(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 2)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                  const __grid_constant__ CUtensorMap map_xl, const __grid_constant__ CUtensorMap map_dyl, const WHParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t rowA = p.Cc * 2, rowB = p.atomB * 2;
    const uint32_t a_bytes = kHaloBW * kHaloBH * rowA;                         // TMA transaction size of the halo tile
    const uint32_t a_stage = (a_bytes + 4 * rowA + 1023) / 1024 * 1024;       // + the junk tap's overhang
    const uint32_t atomB_bytes = 128 * rowB;
    const uint32_t b_bytes = (uint32_t)(p.NT / p.atomB) * atomB_bytes;
    const uint32_t planes = p.dual ? 2u : 1u;                                  // stage = [A_hi][A_lo][B_hi B_lo]
    const uint32_t stage_bytes = planes * a_stage + ((planes * b_bytes + 1023) / 1024) * 1024;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* acc_full = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cset = blockIdx.x % p.csets, range = blockIdx.x / p.csets;       // channel chunk, split-K range
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_img = p.tiles_w * p.tiles_h;
    const int n_my = (p.total_ptiles - range + p.nranges - 1) / p.nranges;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_my; ++i) {
                const int pt = range + i * p.nranges;
                const int b = pt / per_img;
                const int r = pt - b * per_img;
                const int h0 = (r / p.tiles_w) * kHaloTH, w0 = (r % p.tiles_w) * kHaloTW;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], planes * (a_bytes + b_bytes));
                tma_load_4d(sa, &map_x, &full[stage], cset * p.Cc, w0 - 1, h0 - 1, b);
                tma_load_5d(sa + planes * a_stage, &map_dy, &full[stage], 0, w0, h0, b, 0);
                if (p.dual) {
                    tma_load_4d(sa + a_stage, &map_xl, &full[stage], cset * p.Cc, w0 - 1, h0 - 1, b);
                    tma_load_5d(sa + 2 * a_stage + b_bytes, &map_dyl, &full[stage], 0, w0, h0, b, 0);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D=f32, A=B=bf16, A and B MN-major (bits 15, 16), N = NT, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(2 * p.NT >> 3) << 17);
            const uint32_t atoms_per_mma = 128u / (uint32_t)p.Cc;               // 4 or 2 taps (kx) per instruction
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_my; ++i) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t sb = sa + planes * a_stage;
                for (int ky = 0; ky < 3; ++ky)
                    for (int j = 0; j < p.mmas_per_row; ++j) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((ky * p.mmas_per_row + j) * p.NT) * planes;
                        for (int ks = 0; ks < kHaloTH / 2; ++ks) {             // 16 pixels = two patch rows per K step
                            const uint32_t a0 = sa + (uint32_t)(((2 * ks + ky) * kHaloBW) + j * (int)atoms_per_mma) * rowA;
                            const uint64_t da = make_desc_mn_ex(a0, rowA, rowA, kHaloBW * rowA);
                            const uint64_t db = make_desc_mn(sb + (uint32_t)ks * 16 * rowB, rowB, atomB_bytes);
                            const uint32_t acc = (i | ks) ? 1u : 0u;
                            if (!p.dual) {
                                umma_bf16(d_tmem, da, db, idesc, acc);
                                continue;
                            }
                            umma_bf16(d_tmem, da, db, idesc2, acc);           // x_hi * [dy_hi | dy_lo], N = 2*NT <= 256
                            umma_bf16(d_tmem + (uint32_t)p.NT, make_desc_mn_ex(a0 + a_stage, rowA, rowA, kHaloBW * rowA), db, idesc, 1u);
                        }
                    }
                umma_commit(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float* ws = p.ws + (size_t)range * 9 * p.Cin_p * p.Cout_p;
        const int atoms_per_mma = 128 / p.Cc;
        for (int ky = 0; ky < 3; ++ky)
            for (int j = 0; j < p.mmas_per_row; ++j) {
                const int m = q * 32 + lane;                                    // accumulator row = (kx_local, ci)
                const int kx = j * atoms_per_mma + m / p.Cc, ci = cset * p.Cc + m % p.Cc;
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ky * p.mmas_per_row + j) * p.NT) * planes;
                for (int c0 = 0; c0 < p.NT; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t0 + c0, v);
                    if (p.dual) {
                        uint32_t u[16];
                        tmem_ld16(t0 + p.NT + c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
                    } else {
                        tmem_ld_wait();
                    }
                    if (kx < 3 && c0 < p.Cout_p) {
                        float4* dst = reinterpret_cast<float4*>(ws + ((size_t)(ky * 3 + kx) * p.Cin_p + ci) * p.Cout_p + c0);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            dst[e] = n_my > 0 ? make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                                            __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// fills p (except ws); returns the workspace floats, < 0 when the shape is not a halo candidate
static long long plan_wgrad_halo(const ConvGeom& g, WHParams& p, int dual = 0) {
    if (const char* e = getenv("VAE2_WGRAD_HALO")) { if (atoi(e) == 0) return -1; }
    if (g.k != 3 || g.stride != 1 || g.H != g.Ho || g.W != g.Wo) return -1;
    if (!(g.Cin_p == 32 || g.Cin_p % 64 == 0) || g.Cin_p > 512 || g.ldx % 8 || g.Cout_p > 256 || g.Cout_p % 16) return -1;
    p.B = g.B; p.H = g.H; p.W = g.W; p.Cin_p = g.Cin_p; p.Cout_p = g.Cout_p;
    p.Cc = g.Cin_p == 32 ? 32 : 64;
    p.csets = g.Cin_p / p.Cc;
    p.atomB = g.Cout_p % 64 == 0 ? 64 : (g.Cout_p % 32 == 0 ? 32 : 16);
    p.NT = g.Cout_p;
    p.mmas_per_row = p.Cc == 32 ? 1 : 2;
    p.dual = dual ? 1 : 0;
    const int planes = dual ? 2 : 1;
    const int cols = 3 * p.mmas_per_row * p.NT * planes;
    if (cols > 512 || (dual && 2 * p.NT > 256)) return -1;
    p.tmem_cols = next_pow2_cols(cols);
    p.tiles_w = (g.W + kHaloTW - 1) / kHaloTW;
    p.tiles_h = (g.H + kHaloTH - 1) / kHaloTH;
    p.total_ptiles = g.B * p.tiles_w * p.tiles_h;
    const int rowA = p.Cc * 2;
    const long long a_stage = ((long long)kHaloBW * kHaloBH * rowA + 4 * rowA + 1023) / 1024 * 1024;
    const int ctas = p.tmem_cols <= 256 ? 2 : 1;
    const long long stage_b = planes * a_stage + ((long long)planes * p.NT * 128 * 2 + 1023) / 1024 * 1024;
    int stages = (int)((kSmemBudget / ctas - 2048) / stage_b);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return -1;
    p.stages = stages;
    int nranges = ctas * kNumSMs / p.csets;
    if (nranges < 1) nranges = 1;
    if (nranges > p.total_ptiles) nranges = p.total_ptiles;
    p.nranges = nranges;
    return (long long)nranges * 9 * g.Cin_p * g.Cout_p;
}

static int pick_atom(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }

// Fills p (everything except ws) and returns the workspace size in floats; <0 if unsupported.
static long long plan_wgrad_kp(const ConvGeom& g, WParams& p, int kp, int min_stages, int ctas, int dual);

// 64 pixels per stage when that still leaves a 3-deep pipeline, else 32
static long long plan_wgrad(const ConvGeom& g, WParams& p, int dual = 0) {
    int first = 64;
    if (const char* e = getenv("VAE2_WGRAD_KP")) { const int v = atoi(e); if (v == 32) first = 32; }
    // Two co-resident CTAs per SM (each half the shared memory, <= 256 TMEM columns) when the operands are narrow: the
    // kernel is bound by the single issuing thread's barrier round trips (ncu r2a: tensor pipe 7.6 % active, 9 % occupancy
    // on the 18->18 layers), and a second independent issuer per SM overlaps them.  VAE2_WGRAD_CTAS=1 restores one.
    int ctas = 2;
    if (const char* e = getenv("VAE2_WGRAD_CTAS")) { const int v = atoi(e); if (v == 1) ctas = 1; }
    if (ctas == 2) {
        if (first == 64) {
            const long long r = plan_wgrad_kp(g, p, 64, 3, 2, dual);
            if (r >= 0) return r;
        }
        const long long r = plan_wgrad_kp(g, p, 32, 3, 2, dual);
        if (r >= 0) return r;
    }
    if (first == 64) {
        const long long r = plan_wgrad_kp(g, p, 64, 3, 1, dual);
        if (r >= 0) return r;
    }
    return plan_wgrad_kp(g, p, 32, 2, 1, dual);
}

static long long plan_wgrad_kp(const ConvGeom& g, WParams& p, int kp, int min_stages, int ctas, int dual) {
    // K runs over OUTPUT pixels (Ho x Wo); x is sampled at stride g.stride
    p.B = g.B; p.H = g.Ho; p.W = g.Wo; p.sA = g.stride;
    p.n_tiles = (g.Cout_p + 255) / 256;
    p.atomB = pick_atom(g.Cout_p);
    p.NT = (((g.Cout_p + p.n_tiles - 1) / p.n_tiles) + p.atomB - 1) / p.atomB * p.atomB;
    p.Cin_p = g.Cin_p; p.Cout_p = g.Cout_p; p.ksz = g.k; p.taps = g.k * g.k; p.pad = g.k / 2;
    p.atomA = pick_atom(g.Cin_p);
    // pixels per stage: 64 halves the number of stages (each costs the issuing thread ~300 cycles of barrier
    // traffic on top of its MMAs); 32 when the operands are too wide for two 64-pixel stages in shared memory
    int tw = kp;
    while (tw > 8 && tw / 2 >= p.W) tw >>= 1;
    p.TW = tw; p.TH = kp / tw; p.KP = kp;
    p.tiles_w = (p.W + p.TW - 1) / p.TW;
    p.tiles_h = (p.H + p.TH - 1) / p.TH;
    p.total_ptiles = g.B * p.tiles_w * p.tiles_h;
    p.M_total = p.taps * g.Cin_p;
    const int groups_total = (p.M_total + 127) / 128;
    // groups per CTA ("set"): as many as TMEM holds (512 columns), shrunk until the stage ring is deep enough
    const int nchunkA = g.Cin_p / p.atomA;
    const int planes = dual ? 2 : 1;
    p.dual = dual ? 1 : 0;
    int sg_max = (512 / ctas) / (p.NT * planes);
    if (sg_max > groups_total) sg_max = groups_total;
    if (sg_max < 1) return -1;
    if (ctas > 1 && (sg_max < groups_total || p.NT > 64)) return -1;   // co-residency: one set holds every accumulator, narrow N
                                                                      // (measured: 18->18 31.7 vs 35.9 us, 36->36 23.6 vs 27.5 us;
                                                                      //  64->256 1x1 was 12 % slower with halved stages)
    int sg = 0, stages = 0, worst = 0;
    for (int cand = sg_max; cand >= 1; --cand) {
        const int nsets = (groups_total + cand - 1) / cand;
        int w = 0;                      // A atoms per stage: worst case over the sets (whole taps, whole groups)
        for (int s_ = 0; s_ < nsets; ++s_) {
            const int m_lo = s_ * cand * 128;
            const int m_hi = (s_ + 1) * cand * 128;
            const int m_hi_real = m_hi < p.M_total ? m_hi : p.M_total;
            const int t0 = m_lo / g.Cin_p, t1 = (m_hi_real - 1) / g.Cin_p;
            int atoms = (t1 - t0 + 1) * nchunkA;
            const int need = (m_hi - t0 * g.Cin_p + p.atomA - 1) / p.atomA;   // group rows may overhang the loaded taps
            if (need > atoms) atoms = need;
            if (atoms > w) w = atoms;
        }
        const long long a_b = (long long)w * p.KP * p.atomA * 2, b_b = (long long)p.NT * p.KP * 2;
        const long long sb = ((planes * (a_b + b_b) + 1023) / 1024) * 1024;
        int st_ = (int)((kSmemBudget / ctas - (ctas > 1 ? 2048 : 0)) / sb);
        if (st_ > kMaxStages) st_ = kMaxStages;
        if (st_ >= min_stages) { sg = cand; stages = st_; worst = w; break; }
    }
    if (sg == 0) return -1;
    p.set_groups = sg;
    p.nsets = (groups_total + sg - 1) / sg;
    p.tmem_cols = next_pow2_cols(sg * p.NT * planes);
    p.a_atoms_stage = worst;
    p.stages = stages;
    int nranges = ctas * kNumSMs / (p.nsets * p.n_tiles);
    if (nranges < 1) nranges = 1;
    if (nranges > p.total_ptiles) nranges = p.total_ptiles;
    const long long slot = (long long)p.M_total * g.Cout_p;
    const long long cap = (64LL << 20) / 4 / slot;      // <= 64 MB of partials
    if (nranges > cap) nranges = (int)(cap < 1 ? 1 : cap);
    p.nranges = nranges;
    return slot * nranges;
}

}  // namespace tc

int conv_tc_supported(const ConvGeom& g) {
    if (!(g.stride == 1 || (g.stride == 2 && g.k == 3))) return 0;
    if (!(g.k == 1 || g.k == 3)) return 0;
    // lanes in multiples of 16; the one exception are 8-lane OUTPUTS of 1x1 layers (the 270 -> 3 prediction heads write an
    // 8-lane slice of the clip buffer): N is padded to 16 by TMA zero fill and the epilogues store 8 lanes
    const bool head8 = g.Cout_p == 8 && g.k == 1 && g.stride == 1;
    if (g.Cin_p % 16 || (g.Cout_p % 16 && !head8) || g.ldx % 8 || g.ldy % 8) return 0;
    if (head8) { if (const char* e = getenv("VAE2_TC_HEAD8")) { if (atoi(e) == 0) return 0; } }
    if (g.stride == 1 && (g.H != g.Ho || g.W != g.Wo)) return 0;
    if (g.stride == 2 && (g.Ho != (g.H + 1) / 2 || g.Wo != (g.W + 1) / 2)) return 0;
    return tc::encode_fn() != nullptr ? 1 : 0;
}

namespace tc {
struct Launch {
    const void* a; int lda, Ck, AH, AW, a_estride;   // gathered tensor [B][AH][AW][lda], Ck lanes, TMA element stride
    const void* wq; int Cn, wtaps;                    // packed bf16 weights [wtaps][Cn][Ck]
    void* out; int ldo, OH, OW;
    const float* bias;
    int B, MH, MW, sO, oh_off, ow_off, sA;
    int ntaps; signed char dh[9], dw[9], wt[9];
    int accumulate;
};
}  // namespace tc

static int launch_tc(const tc::Launch& L, cudaStream_t st) {
    using namespace tc;
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return VAE2_ERR_UNSUPPORTED;
    if (L.MH <= 0 || L.MW <= 0) return VAE2_OK;
    Params p;
    p.B = L.B; p.MH = L.MH; p.MW = L.MW; p.OH = L.OH; p.OW = L.OW;
    p.sO = L.sO; p.oh_off = L.oh_off; p.ow_off = L.ow_off; p.sA = L.sA;
    p.Cn = L.Cn; p.ldo = L.ldo;
    p.taps = L.ntaps;
    for (int i = 0; i < 9; ++i) { p.tap_dh[i] = L.dh[i]; p.tap_dw[i] = L.dw[i]; p.tap_w[i] = L.wt[i]; }
    p.KC = pick_kc(L.Ck);
    p.kchunks = (L.Ck + p.KC - 1) / p.KC;
    p.n_tiles = (L.Cn + 255) / 256;
    p.NT = (((L.Cn + p.n_tiles - 1) / p.n_tiles) + 15) / 16 * 16;
    p.accumulate = L.accumulate;
    p.bias = L.bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(L.out);
    // ---- halo path: 3x3, unit strides, all taps' weights resident in shared memory ----
    {
        bool halo = L.ntaps == 9 && L.sA == 1 && L.a_estride == 1 && L.sO == 1 && L.oh_off == 0 && L.ow_off == 0 &&
                    p.n_tiles == 1 && 2 * p.NT <= 512 && L.Cn % 16 == 0;
        for (int i = 0; halo && i < 9; ++i) halo = L.dh[i] >= -1 && L.dh[i] <= 1 && L.dw[i] >= -1 && L.dw[i] <= 1;
        if (const char* e = getenv("VAE2_TC_HALO")) { if (atoi(e) == 0) halo = false; }
        // K chunk: a halo stage costs one barrier round trip per chunk (not per tap), so the MMA count decides:
        // fewest 16-channel steps, ties to the wider chunk (fewer TMA boxes)
        int hkc = 16, best = 1 << 30;
        for (int kc = 16; kc <= 64; kc *= 2) {
            const int cost = ((L.Ck + kc - 1) / kc) * (kc / 16) * 8 + (L.Ck + kc - 1) / kc;
            if (cost <= best) { best = cost; hkc = kc; }
        }
        const int hchunks = (L.Ck + hkc - 1) / hkc;
        const int row_bytes = hkc * 2;
        const int atom = 8 * row_bytes;                                      // swizzle repeat: 8 rows of one span
        const int a_stage = (kHaloBW * kHaloBH * row_bytes + 1023) / 1024 * 1024;
        const int w_block = (p.NT * row_bytes + atom - 1) / atom * atom;
        const int w_bytes = (hchunks * 9 * w_block + 1023) / 1024 * 1024;
        const int bar_bytes = 1024 /*align slack*/ + (2 * kMaxStages + 5) * 8 + 16;
        // as many co-resident CTAs (independent MMA issuers) as shared memory and TMEM allow; a single issuer per SM
        // loses to conv_tc_kernel's two, so the halo path needs at least 2
        int ctas = 4;
        while (ctas > 1 && (w_bytes + 2 * a_stage + bar_bytes > 227 * 1024 / ctas - 1024 ||
                            next_pow2_cols(2 * p.NT) * ctas > 512)) --ctas;
        if (halo && ctas >= 2) {
            p.KC = hkc; p.kchunks = hchunks;
            int stages = (227 * 1024 / ctas - 1024 - w_bytes - bar_bytes) / a_stage;
            if (stages > kMaxStages) stages = kMaxStages;
            if (stages >= 2) {
                p.stages = stages;
                p.acc_stages = 2;
                p.tmem_cols = next_pow2_cols(2 * p.NT);
                p.TW = kHaloTW; p.TH = kHaloTH;
                p.tiles_w = (L.MW + kHaloTW - 1) / kHaloTW;
                p.tiles_h = (L.MH + kHaloTH - 1) / kHaloTH;
                p.total_tiles = L.B * p.tiles_w * p.tiles_h;
                CUtensorMap map_a, map_w;
                {
                    cuuint64_t dims[4] = {(cuuint64_t)L.Ck, (cuuint64_t)L.AW, (cuuint64_t)L.AH, (cuuint64_t)L.B};
                    cuuint64_t strides[3] = {(cuuint64_t)L.lda * 2, (cuuint64_t)L.AW * L.lda * 2, (cuuint64_t)L.AH * L.AW * L.lda * 2};
                    cuuint32_t box[4] = {(cuuint32_t)p.KC, (cuuint32_t)kHaloBW, (cuuint32_t)kHaloBH, 1};
                    cuuint32_t es[4] = {1, 1, 1, 1};
                    if (enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(L.a), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                        halo = false;
                }
                if (halo) {
                    cuuint64_t dims[3] = {(cuuint64_t)L.Ck, (cuuint64_t)L.Cn, (cuuint64_t)L.wtaps};
                    cuuint64_t strides[2] = {(cuuint64_t)L.Ck * 2, (cuuint64_t)L.Cn * L.Ck * 2};
                    cuuint32_t box[3] = {(cuuint32_t)p.KC, (cuuint32_t)p.NT, 1};
                    cuuint32_t es[3] = {1, 1, 1};
                    if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(L.wq), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                        halo = false;
                }
                const size_t smem = (size_t)w_bytes + (size_t)stages * a_stage + bar_bytes;
                static bool attr_set_h = false;
                if (!attr_set_h) {
                    if (cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                        return VAE2_ERR_CUDA;
                    attr_set_h = true;
                }
                if (halo) {
                    const int grid = p.total_tiles < ctas * kNumSMs ? p.total_tiles : ctas * kNumSMs;
                    note_kernel("tc::conv_tc_halo_kernel");
                    conv_tc_halo_kernel<<<grid, kThreads, smem, st>>>(map_a, map_w, p);
                    return check_launch();
                }
            }
        }
    }
    p.KC = pick_kc(L.Ck);                          // (the halo attempt may have changed the chunking)
    p.kchunks = (L.Ck + p.KC - 1) / p.KC;
    int tw = 128;
    while (tw > 8 && tw / 2 >= L.MW) tw >>= 1;
    p.TW = tw; p.TH = 128 / tw;
    p.tiles_w = (L.MW + p.TW - 1) / p.TW;
    p.tiles_h = (L.MH + p.TH - 1) / p.TH;
    p.total_tiles = L.B * p.tiles_w * p.tiles_h * p.n_tiles;
    const int row_bytes = p.KC * 2;
    const int a_bytes = 128 * row_bytes;
    const int b_bytes = ((p.NT * row_bytes + 1023) / 1024) * 1024;
    // Two co-resident CTAs per SM (each with half the shared memory and <= 256 TMEM columns) when the tiles are
    // small: tcgen05.mma issue costs the issuing thread ~110 cycles regardless of N, so narrow-N convs are
    // issue-bound and a second, independent issuer per SM doubles their throughput.
    int ctas = 1;
    if (p.NT <= 32 && 4 * (a_bytes + b_bytes) + 2048 <= kSmemBudget / 4) ctas = 4;
    else if (p.NT <= 128 && 4 * (a_bytes + b_bytes) + 2048 <= kSmemBudget / 2) ctas = 2;
    if (const char* e = getenv("VAE2_TC_CTAS")) { const int v = atoi(e); if (v >= 1 && v < ctas) ctas = v; }
    const int budget = kSmemBudget / ctas;
    int stages = budget / (a_bytes + b_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return VAE2_ERR_UNSUPPORTED;
    p.stages = stages;
    p.acc_stages = (2 * p.NT <= 512 / ctas) ? 2 : 1;
    p.tmem_cols = next_pow2_cols(p.acc_stages * p.NT);
    p.accumulate = L.accumulate;
    p.bias = L.bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(L.out);

    CUtensorMap map_a, map_w;
    {
        const cuuint32_t es_ = (cuuint32_t)L.a_estride;
        cuuint64_t dims[4] = {(cuuint64_t)L.Ck, (cuuint64_t)L.AW, (cuuint64_t)L.AH, (cuuint64_t)L.B};
        cuuint64_t strides[3] = {(cuuint64_t)L.lda * 2, (cuuint64_t)L.AW * L.lda * 2, (cuuint64_t)L.AH * L.AW * L.lda * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.KC, (cuuint32_t)p.TW * es_, (cuuint32_t)p.TH * es_, 1};
        cuuint32_t es[4] = {1, es_, es_, 1};
        if (enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(L.a), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return VAE2_ERR_ARG;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)L.Ck, (cuuint64_t)L.Cn, (cuuint64_t)L.wtaps};
        cuuint64_t strides[2] = {(cuuint64_t)L.Ck * 2, (cuuint64_t)L.Cn * L.Ck * 2};
        cuuint32_t box[3] = {(cuuint32_t)p.KC, (cuuint32_t)p.NT, 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(L.wq), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return VAE2_ERR_ARG;
    }
    const size_t smem = (size_t)stages * (a_bytes + b_bytes) + 1024 /*align slack*/ + (2 * kMaxStages + 4) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return VAE2_ERR_CUDA;
        attr_set = true;
    }
    int grid = p.total_tiles < ctas * kNumSMs ? p.total_tiles : ctas * kNumSMs;
    note_kernel("tc::conv_tc_kernel");
    conv_tc_kernel<<<grid, kThreads, smem, st>>>(map_a, map_w, p);
    return check_launch();
}

int conv_fwd_tc(const void* x, const void* wq, const float* bias, void* y, const ConvGeom& g, float*, cudaStream_t st) {
    if (!conv_tc_supported(g)) return VAE2_ERR_UNSUPPORTED;
    tc::Launch L{};
    L.a = x; L.lda = g.ldx; L.Ck = g.Cin_p; L.AH = g.H; L.AW = g.W; L.a_estride = g.stride;
    L.wq = wq; L.Cn = g.Cout_p; L.wtaps = g.k * g.k;
    L.out = y; L.ldo = g.ldy; L.OH = g.Ho; L.OW = g.Wo; L.bias = bias;
    L.B = g.B; L.MH = g.Ho; L.MW = g.Wo; L.sO = 1; L.oh_off = 0; L.ow_off = 0; L.sA = g.stride;
    L.ntaps = g.k * g.k;
    for (int t = 0; t < L.ntaps; ++t) { L.dh[t] = (signed char)(t / g.k - g.pad); L.dw[t] = (signed char)(t % g.k - g.pad); L.wt[t] = (signed char)t; }
    L.accumulate = 0;
    return launch_tc(L, st);
}

// dx (=|+=) conv_transpose(dy): A = dy, weights = wqT [tap][Cin_p][Cout_p].
// stride 1: one launch with mirrored tap offsets.  stride 2: dx pixels split into the four parity
// classes (h%2, w%2); class (ph, pw) only sees taps with ky = ph+1 (mod 2), kx = pw+1 (mod 2) and is a
// small stride-1 conv over dy whose results land on the class's strided pixel grid.
int conv_dgrad_tc(const void* dy, const void* wqT, void* dx, const ConvGeom& g, int accumulate, cudaStream_t st) {
    if (!conv_tc_supported(g)) return VAE2_ERR_UNSUPPORTED;
    tc::Launch L{};
    L.a = dy; L.lda = g.ldy; L.Ck = g.Cout_p; L.AH = g.Ho; L.AW = g.Wo; L.a_estride = 1;
    L.wq = wqT; L.Cn = g.Cin_p; L.wtaps = g.k * g.k;
    L.out = dx; L.ldo = g.ldx; L.OH = g.H; L.OW = g.W; L.bias = nullptr;
    L.B = g.B; L.sA = 1; L.accumulate = accumulate;
    if (g.stride == 1) {
        L.MH = g.H; L.MW = g.W; L.sO = 1; L.oh_off = 0; L.ow_off = 0;
        L.ntaps = g.k * g.k;
        for (int t = 0; t < L.ntaps; ++t) { L.dh[t] = (signed char)(g.pad - t / g.k); L.dw[t] = (signed char)(g.pad - t % g.k); L.wt[t] = (signed char)t; }
        return launch_tc(L, st);
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
            if (int e = launch_tc(L, st)) return e;
        }
    return VAE2_OK;
}

}  // namespace vae2

namespace vae2 {


static int make_map5(tc::EncodeTiledFn enc, CUtensorMap* m, const void* base, int atom, int C, int boxC, int ld, int B, int H,
                     int W, int TW, int TH, int estride) {
    const cuuint32_t e = (cuuint32_t)estride;
    // C < atom (8-lane head outputs): the inner dimension is the real width, the box still spans `atom` lanes (zero fill)
    cuuint64_t dims[5] = {(cuuint64_t)(C < atom ? C : atom), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B,
                          (cuuint64_t)((C + atom - 1) / atom)};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2, (cuuint64_t)atom * 2};
    cuuint32_t box[5] = {(cuuint32_t)atom, (cuuint32_t)TW * e, (cuuint32_t)TH * e, 1, (cuuint32_t)(boxC / atom)};
    cuuint32_t es[5] = {1, e, e, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swz(atom), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

namespace tc {
// one product's split-K partials: which kernel, how many slots, floats per slot
struct WgradPlan {
    bool halo;
    WParams w;
    WHParams h;
    long long part;     // floats of all slots
    int nslots;
    long long n;        // floats of one slot = taps * Cin_p * Cout_p
};

static int plan_wgrad_any(const ConvGeom& g, WgradPlan& P, int dual = 0) {
    P.part = plan_wgrad_halo(g, P.h, dual);
    if (P.part > 0) {
        P.halo = true; P.nslots = P.h.nranges; P.n = 9LL * g.Cin_p * g.Cout_p;
        return 0;
    }
    P.halo = false;
    P.part = plan_wgrad(g, P.w, dual);
    if (P.part < 0) return -1;
    P.nslots = P.w.nranges; P.n = (long long)P.w.M_total * g.Cout_p;
    return 0;
}

// ws[slot][tap][Cin_p][Cout_p] = partial weight gradients of bf16 x [B][H][W][ldx] and dy [B][Ho][Wo][ldy]
// (xl, dyl: the low planes of a dual plan, else null)
static int wgrad_partials(const void* x, const void* dy, float* ws, const ConvGeom& g, const WgradPlan& P, cudaStream_t st,
                          const void* xl = nullptr, const void* dyl = nullptr) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return VAE2_ERR_UNSUPPORTED;
    CUtensorMap map_x, map_dy, map_xl, map_dyl;
    const bool dual = P.halo ? P.h.dual != 0 : P.w.dual != 0;
    if (dual && (xl == nullptr || dyl == nullptr)) return VAE2_ERR_ARG;
    if (P.halo) {
        WHParams p = P.h;
        p.ws = ws;
        for (int pl = 0; pl < (dual ? 2 : 1); ++pl) {
            cuuint64_t dims[4] = {(cuuint64_t)g.Cin_p, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
            cuuint64_t strides[3] = {(cuuint64_t)g.ldx * 2, (cuuint64_t)g.W * g.ldx * 2, (cuuint64_t)g.H * g.W * g.ldx * 2};
            cuuint32_t box[4] = {(cuuint32_t)p.Cc, (cuuint32_t)kHaloBW, (cuuint32_t)kHaloBH, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            if (enc(pl ? &map_xl : &map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pl ? xl : x), dims, strides, box,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.Cc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return VAE2_ERR_ARG;
        }
        if (make_map5(enc, &map_dy, dy, p.atomB, g.Cout_p, p.NT, g.ldy, g.B, g.Ho, g.Wo, kHaloTW, kHaloTH, 1)) return VAE2_ERR_ARG;
        if (dual) {
            if (make_map5(enc, &map_dyl, dyl, p.atomB, g.Cout_p, p.NT, g.ldy, g.B, g.Ho, g.Wo, kHaloTW, kHaloTH, 1)) return VAE2_ERR_ARG;
        } else {
            map_xl = map_x; map_dyl = map_dy;
        }
        const int rowA = p.Cc * 2, planes = dual ? 2 : 1;
        const long long a_stage = ((long long)kHaloBW * kHaloBH * rowA + 4 * rowA + 1023) / 1024 * 1024;
        const long long b_stage = ((long long)planes * p.NT * 128 * 2 + 1023) / 1024 * 1024;
        const size_t smem = (size_t)p.stages * (planes * a_stage + b_stage) + 1024 + (2 * kMaxStages + 4) * 8 + 16;
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                return VAE2_ERR_CUDA;
            attr_set = true;
        }
        wgrad_halo_kernel<<<p.nranges * p.csets, kThreads, smem, st>>>(map_x, map_dy, map_xl, map_dyl, p);
        return check_launch();
    }
    WParams p = P.w;
    p.ws = ws;
    if (make_map5(enc, &map_x, x, p.atomA, g.Cin_p, g.Cin_p, g.ldx, g.B, g.H, g.W, p.TW, p.TH, g.stride)) return VAE2_ERR_ARG;
    if (make_map5(enc, &map_dy, dy, p.atomB, g.Cout_p, p.NT, g.ldy, g.B, g.Ho, g.Wo, p.TW, p.TH, 1)) return VAE2_ERR_ARG;
    if (dual) {
        if (make_map5(enc, &map_xl, xl, p.atomA, g.Cin_p, g.Cin_p, g.ldx, g.B, g.H, g.W, p.TW, p.TH, g.stride)) return VAE2_ERR_ARG;
        if (make_map5(enc, &map_dyl, dyl, p.atomB, g.Cout_p, p.NT, g.ldy, g.B, g.Ho, g.Wo, p.TW, p.TH, 1)) return VAE2_ERR_ARG;
    } else {
        map_xl = map_x; map_dyl = map_dy;
    }
    const long long a_bytes = (long long)p.a_atoms_stage * p.KP * p.atomA * 2;
    const long long b_bytes = (long long)p.NT * p.KP * 2;
    const long long stage_bytes = (((dual ? 2 : 1) * (a_bytes + b_bytes) + 1023) / 1024) * 1024;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (2 * kMaxStages + 4) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return VAE2_ERR_CUDA;
        attr_set = true;
    }
    wgrad_tc_kernel<<<p.nsets * p.n_tiles * p.nranges, kThreads, smem, st>>>(map_x, map_dy, map_xl, map_dyl, p);
    return check_launch();
}
}  // namespace tc

long long conv_wgrad_tc_workspace(const ConvGeom& g) {
    if (!conv_tc_supported(g)) return -1;
    tc::WgradPlan P;
    return tc::plan_wgrad_any(g, P) ? -1 : P.part;
}

// dwp [tap][Cin_p][Cout_p] fp32 (overwritten); ws: at least conv_wgrad_tc_workspace(g) floats
int conv_wgrad_tc(const void* x, const void* dy, float* dwp, float* ws, const ConvGeom& g, cudaStream_t st) {
    using namespace tc;
    if (!conv_tc_supported(g)) return VAE2_ERR_UNSUPPORTED;
    WgradPlan P;
    if (plan_wgrad_any(g, P)) return VAE2_ERR_UNSUPPORTED;
    note_kernel(P.halo ? "tc::wgrad_halo_kernel" : "tc::wgrad_tc_kernel");
    if (int e = wgrad_partials(x, dy, ws, g, P, st)) return e;
    launch_wgrad_reduce(ws, dwp, P.n, P.nslots, st);
    return check_launch();
}

// =================================================================================================
// fp32 weight gradient on tcgen05 ("f32x2"): fp32 activations are split into two bf16 planes
// (hi = bf16(x), lo = bf16(x - hi); what is dropped is <= 2^-17 |x|) and the gradient is the sum of three plane
// products  hi*hi + hi*lo + lo*hi, each one launch of wgrad_tc_kernel above with its OWN accumulators and split-K
// slots (so the small products never share a truncating accumulator with the leading one); the slots of all three
// are folded by wgrad_reduce_kernel in fixed order and cropped into the fp32 path's [tap][Cin_p][Cout_p] layout.
// Replaces conv_igemm's split-K FMA weight gradient for the >= 40-lane layers of the fp32 path
// (ncu r2b: 270->270 1x1 @B=4 3577 us, 64->64 3x3 1279 us -- 27 % of the fp32 step).
// =================================================================================================
namespace tc {

// src fp32 [npix][ld] (lanes used: C) -> hi, lo bf16 [npix][Cq] (Cq = C padded to 16, pad lanes zero)
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                    long long npix, int C, int ld, int Cq) {
    const int vec = Cq / 8;                                   // 8 output lanes (16 bytes of bf16) per thread
    const long long total = npix * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / vec;
        const int c0 = (int)(i - p * vec) * 8;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k += 4) {
            if (c0 + k < C) {                                 // C and ld are multiples of 4
                const float4 f = *reinterpret_cast<const float4*>(src + p * ld + c0 + k);
                v[k] = f.x; v[k + 1] = f.y; v[k + 2] = f.z; v[k + 3] = f.w;
            } else {
                v[k] = v[k + 1] = v[k + 2] = v[k + 3] = 0.f;
            }
        }
        uint4 oh, ol;
        __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&oh);
        __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(&ol);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            h[k] = __float2bfloat16_rn(v[k]);
            l[k] = __float2bfloat16_rn(v[k] - __bfloat162float(h[k]));
        }
        *reinterpret_cast<uint4*>(hi + p * Cq + c0) = oh;
        *reinterpret_cast<uint4*>(lo + p * Cq + c0) = ol;
    }
}

// dwp[t][ci][co] (Cin_p x Cout_p lanes) = src[t][ci][co] (Ciq x Coq lanes)
__global__ void crop_dw_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int Cin_p, int Cout_p,
                               int Ciq, int Coq) {
    const int total = taps * Cin_p * Cout_p;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout_p;
        const int ci = (i / Cout_p) % Cin_p;
        const int t = i / (Cout_p * Cin_p);
        dst[i] = src[((long long)t * Ciq + ci) * Coq + co];
    }
}

static ConvGeom geom16(const ConvGeom& g) {
    ConvGeom q = g;
    q.Cin_p = (g.Cin_p + 15) / 16 * 16; q.ldx = q.Cin_p;
    q.Cout_p = (g.Cout_p + 15) / 16 * 16; q.ldy = q.Cout_p;
    return q;
}

// plan of the plane products: the single-pass dual plan when it exists (VAE2_WGRAD_DUAL=0: three separate products)
static int f32x2_plan(const ConvGeom& q, WgradPlan& P) {
    bool dual = true;
    if (const char* e = getenv("VAE2_WGRAD_DUAL")) { if (atoi(e) == 0) dual = false; }
    WgradPlan D;
    if (dual && plan_wgrad_any(q, D, 1) == 0) {
        // the second accumulator halves the M groups a CTA holds; when that splits the problem into more sets (each re-reads
        // the operand tiles: 270 -> 270 measured 1.04 vs 0.87 ms) the three separate products win
        if (D.halo || plan_wgrad_any(q, P, 0) != 0 || P.halo || D.w.nsets <= P.w.nsets) { P = D; return 0; }
        return 0;
    }
    return plan_wgrad_any(q, P, 0);
}

}  // namespace tc

// bytes of workspace conv_wgrad_f32x2 needs (negative: unsupported geometry)
long long conv_wgrad_f32x2_workspace(const ConvGeom& g) {
    const ConvGeom q = tc::geom16(g);
    if (!conv_tc_supported(q)) return -1;
    tc::WgradPlan P;
    if (tc::f32x2_plan(q, P)) return -1;
    const long long part = P.part;                          // floats of ONE product's split-K slots (dual plans: of the pass)
    const long long nx = (long long)g.B * g.H * g.W * q.Cin_p, ny = (long long)g.B * g.Ho * g.Wo * q.Cout_p;
    const long long dw16 = P.n;
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    return al(2 * nx * 2) + al(2 * ny * 2) + al(3 * part * 4) + al(dw16 * 4);
}

int conv_wgrad_f32x2(const float* x, const float* dy, float* dwp, void* workspace, const ConvGeom& g, cudaStream_t st) {
    using namespace tc;
    const ConvGeom q = geom16(g);
    if (!conv_tc_supported(q) || g.Cin_p % 4 || g.Cout_p % 4 || g.ldx % 4 || g.ldy % 4) return VAE2_ERR_UNSUPPORTED;
    WgradPlan P;
    if (f32x2_plan(q, P)) return VAE2_ERR_UNSUPPORTED;
    const bool dual = P.halo ? P.h.dual != 0 : P.w.dual != 0;
    const long long part = P.part;
    const long long npx = (long long)g.B * g.H * g.W, npy = (long long)g.B * g.Ho * g.Wo;
    const long long nx = npx * q.Cin_p, ny = npy * q.Cout_p;
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(w);
    __nv_bfloat16* xl = xh + nx;
    w += al(2 * nx * 2);
    __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(w);
    __nv_bfloat16* yl = yh + ny;
    w += al(2 * ny * 2);
    float* slots = reinterpret_cast<float*>(w);
    w += al(3 * part * 4);
    float* dw16 = reinterpret_cast<float*>(w);
    note_kernel(P.halo ? (dual ? "tc::wgrad_halo_kernel dual (f32x2 planes)" : "tc::wgrad_halo_kernel (f32x2 planes)")
                       : (dual ? "tc::wgrad_tc_kernel dual (f32x2 planes)" : "tc::wgrad_tc_kernel (f32x2 planes)"));
    split_planes_kernel<<<stream_grid(npx * (q.Cin_p / 8), 256), 256, 0, st>>>(x, xh, xl, npx, g.Cin_p, g.ldx, q.Cin_p);
    split_planes_kernel<<<stream_grid(npy * (q.Cout_p / 8), 256), 256, 0, st>>>(dy, yh, yl, npy, g.Cout_p, g.ldy, q.Cout_p);
    if (int e = check_launch()) return e;
    const long long n = P.n;
    if (dual) {
        // one pass: x_hi*dy_hi in one accumulator, x_hi*dy_lo + x_lo*dy_hi in a second one, added in the epilogue
        if (int e = wgrad_partials(xh, yh, slots, q, P, st, xl, yl)) return e;
        launch_wgrad_reduce(slots, dw16, n, P.nslots, st);
    } else {
        const __nv_bfloat16* xa[3] = {xl, xh, xh};      // smallest products first in the fold: lo*hi, hi*lo, hi*hi
        const __nv_bfloat16* ya[3] = {yh, yl, yh};
        for (int t = 0; t < 3; ++t)
            if (int e = wgrad_partials(xa[t], ya[t], slots + (long long)t * part, q, P, st)) return e;
        launch_wgrad_reduce(slots, dw16, n, 3 * P.nslots, st);
    }
    const int total = g.k * g.k * g.Cin_p * g.Cout_p;
    crop_dw_kernel<<<(total + 255) / 256, 256, 0, st>>>(dw16, dwp, g.k * g.k, g.Cin_p, g.Cout_p, q.Cin_p, q.Cout_p);
    return check_launch();
}

}  // namespace vae2
