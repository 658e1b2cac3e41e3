// Register-tiled direct convolution on the CUDA cores: the exact-fp32 forward pass and stride-1 data gradient of
// the NARROW layers (HRNet branches: 18/36/72 channels, i.e. 20/36/72 fp32 lanes).
//
// Same contract as conv_simt.cu's implicit GEMM (which stays for wide layers, the weight gradient, the stride-2
// data gradient and bf16 storage), i.e. the cuDNN calls behind the reference's nn.Conv2d sites
// (lib/models/enc_hrnet.py:27-30, 38-41, 70-76, 188-218, 381-404), fp32 FMA accumulate.
//
// Why not the GEMM tiling here: a 32- or 64-column GEMM tile pads 20/36/72 lanes by up to 78 %, a 16-deep K chunk
// splits a 20-channel tap into 16 + 4, and every chunk pays two block barriers plus a transposing shared-memory
// store of an im2col tile that repeats every input pixel nine times.  Here
//   * the CTA stages its input patch (output patch + halo, all input channels) in shared memory ONCE, pixel-major
//     with a pitch of 4*odd words, so that 8 consecutive pixels' float4 hit 32 distinct banks;
//   * a thread owns TM = 4 output pixels (same column, 4 consecutive rows: a warp is 32 columns wide) x TN output
//     channels, TN dividing the lanes exactly (20 -> 20, 36 -> 12, 72 -> 24), accumulators in registers;
//   * the weights of one (tap, channel block) sit in shared memory as [ci][co] -- the packed layout, so the copy
//     is a linear cp.async -- and are read as warp-broadcast float4; the next block streams in behind the math.
// Per 4 input channels a thread issues 4 conflict-free + TN broadcast shared float4 loads for 16*TN FMAs; the only
// block barriers are the two per weight block.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vae2 {
namespace direct {

constexpr int TM = 4;

struct Params {
    const float* a;          // gathered tensor [B][IH][IW][lda]
    const float* w;          // packed weights [tap][CIN][COUT]
    const float* bias;       // [COUT] or null
    float* out;              // [B][OH][OW][ldo]
    int B, IH, IW, lda, OH, OW, ldo;
    int CIN, COUT;           // K lanes per tap (multiple of 4); output lanes = TN * nsplit
    int taps, stride;
    int dh[9], dw[9], wt[9]; // grid point (i, j) under tap t reads input pixel (i*stride + dh[t], j*stride + dw[t])
                             // and weight slice wt[t]
    int dh_min, dw_min;
    int MH, MW;              // grid of output points handled by this launch
    int sO, oh_off, ow_off;  // grid point (i, j) is output pixel (i*sO + oh_off, j*sO + ow_off)
    int nsplit, groups;      // threads = groups * nsplit; thread = (channel slice, pixel group)
    int PW, PH;              // output patch: PW columns x PH rows (PH = TM * row groups)
    int PWin, PHin, pitch;   // staged input patch (pixels) and its pixel pitch in floats (4 * odd)
    int tiles_w, tiles_h;
    int kblk, nblk;          // channels per staged weight block, blocks per tap
    int accumulate;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

template <int TN>
__global__ void __launch_bounds__(192, 2)      // <= 170 registers: three 128-thread CTAs (12 warps) per SM
conv_direct_kernel(const Params p) {
    extern __shared__ __align__(16) float smem[];
    float* asm_ = smem;                                        // [PHin*PWin][pitch]
    float* wsm = smem + (size_t)p.PHin * p.PWin * p.pitch;     // [2][kblk * COUT]
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int ns = tid / p.groups, grp = tid - ns * p.groups;  // a warp shares one channel slice
    const int col = grp % p.PW, rowg = grp / p.PW;
    const int per_img = p.tiles_w * p.tiles_h;
    const int b = blockIdx.x / per_img;
    const int r = blockIdx.x - b * per_img;
    const int oh0 = (r / p.tiles_w) * p.PH, ow0 = (r % p.tiles_w) * p.PW;
    const int n0 = ns * TN;
    const int wstage = p.kblk * p.COUT;

    // ---- stage the input patch (zero outside the image: that IS the padding) ----
    {
        const int q4 = p.CIN / 4;
        const int ih0 = oh0 * p.stride + p.dh_min, iw0 = ow0 * p.stride + p.dw_min;
        const int nvec = p.PHin * p.PWin * q4;
        for (int v = tid; v < nvec; v += nthreads) {
            const int pix = v / q4, q = v - pix * q4;
            const int pr = pix / p.PWin, pc = pix - pr * p.PWin;
            const int ih = ih0 + pr, iw = iw0 + pc;
            float* dst = asm_ + (size_t)pix * p.pitch + 4 * q;
            if (ih >= 0 && ih < p.IH && iw >= 0 && iw < p.IW)
                cp_async16(dst, p.a + (((long long)b * p.IH + ih) * p.IW + iw) * p.lda + 4 * q);
            else
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const int nstage = p.taps * p.nblk;
    auto stage_load = [&](int s) {
        const int tap = s / p.nblk, cb = s - tap * p.nblk;
        const int c0 = cb * p.kblk;
        const int kb = min(p.kblk, p.CIN - c0);
        const float4* src = reinterpret_cast<const float4*>(p.w + ((long long)p.wt[tap] * p.CIN + c0) * p.COUT);
        float4* dst = reinterpret_cast<float4*>(wsm + (s & 1) * wstage);
        const int nvec = kb * p.COUT / 4;
        for (int v = tid; v < nvec; v += nthreads) cp_async16(dst + v, src + v);
    };
    stage_load(0);
    cp_async_commit();                     // group 0 = patch + weight block 0

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const float* arow[TM];
    for (int s = 0; s < nstage; ++s) {
        if (s + 1 < nstage) stage_load(s + 1);
        cp_async_commit();
        const int tap = s / p.nblk, cb = s - tap * p.nblk;
        const int c0 = cb * p.kblk;
        const int kb = min(p.kblk, p.CIN - c0);
        {
            const int pc = col * p.stride + p.dw[tap] - p.dw_min;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int pr = (rowg * TM + i) * p.stride + p.dh[tap] - p.dh_min;
                arow[i] = asm_ + (size_t)(pr * p.PWin + pc) * p.pitch + c0;
            }
        }
        cp_async_wait<1>();
        __syncthreads();                 // patch and weight block s have landed for every thread
        const float* wb = wsm + (s & 1) * wstage + n0;
        // software pipeline: the weight row of step k+1 (and the next activation quad) are fetched before the
        // FMAs of step k, so the shared-memory latency hides behind 4*TN FMAs even with few resident warps
        float4 av[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(arow[i]);
        float bv[TN], bn[TN];
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(wb + j);
            bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
        }
#pragma unroll 1
        for (int c = 0; c < kb; c += 4) {
            float4 an[TM];
            const int cn = c + 4 < kb ? c + 4 : c;          // (the last prefetch re-reads the current quad: harmless)
#pragma unroll
            for (int i = 0; i < TM; ++i) an[i] = *reinterpret_cast<const float4*>(arow[i] + cn);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int kn = (c + kk + 1 < kb) ? c + kk + 1 : c + kk;
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(wb + kn * p.COUT + j);
                    bn[j] = t.x; bn[j + 1] = t.y; bn[j + 2] = t.z; bn[j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float a = comp(av[i], kk);
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a, bv[j], acc[i][j]);
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = bn[j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = an[i];
        }
        __syncthreads();                 // everyone is done with buffer s&1 before block s+2 overwrites it
    }

    const int gj = ow0 + col;
    if (gj >= p.MW) return;
    const int ow = gj * p.sO + p.ow_off;
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = p.bias != nullptr ? p.bias[n0 + j] : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gi = oh0 + rowg * TM + i;
        if (gi >= p.MH) continue;
        const int oh = gi * p.sO + p.oh_off;
        float* o = p.out + (((long long)b * p.OH + oh) * p.OW + ow) * p.ldo + n0;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            float4 v = make_float4(acc[i][j] + bv[j], acc[i][j + 1] + bv[j + 1], acc[i][j + 2] + bv[j + 2], acc[i][j + 3] + bv[j + 3]);
            if (p.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(o + j);
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            *reinterpret_cast<float4*>(o + j) = v;
        }
    }
}

template <int TN>
static int launch_t(const Params& p, size_t smem, int grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_direct_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return VAE2_ERR_CUDA;
        attr = true;
    }
    note_kernel("direct::conv_direct_kernel");
    conv_direct_kernel<TN><<<grid, p.groups * p.nsplit, smem, st>>>(p);
    return check_launch();
}

static int launch(Params& p, cudaStream_t st) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("VAE2_DIRECT"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled) return VAE2_ERR_UNSUPPORTED;
    // narrow layers only: wider ones amortise the GEMM tiling and would re-stage too many weights per pixel here
    if (p.CIN > 72 || p.COUT > 72 || p.CIN % 4) return VAE2_ERR_UNSUPPORTED;
    if (p.COUT % 64 == 0 || p.CIN % 64 == 0) return VAE2_ERR_UNSUPPORTED;    // measured: the 128x64 GEMM tile wins there
    int tn = 0;
    for (int cand : {24, 20, 16, 12, 8, 4})
        if (p.COUT % cand == 0) { tn = cand; break; }
    if (tn == 0) return VAE2_ERR_UNSUPPORTED;
    p.nsplit = p.COUT / tn;
    int dh_max = p.dh[0], dw_max = p.dw[0];
    p.dh_min = p.dh[0]; p.dw_min = p.dw[0];
    for (int t = 1; t < p.taps; ++t) {
        if (p.dh[t] < p.dh_min) p.dh_min = p.dh[t];
        if (p.dw[t] < p.dw_min) p.dw_min = p.dw[t];
        if (p.dh[t] > dh_max) dh_max = p.dh[t];
        if (p.dw[t] > dw_max) dw_max = p.dw[t];
    }
    // patch: a warp is PW = 32 columns (fewer for narrow images) x TM rows; row groups until ~128-192 threads
    int pw = 32;
    while (pw > 4 && pw / 2 >= p.MW) pw >>= 1;
    p.PW = pw;
    int rowgroups = 1;
    while (pw * rowgroups * 2 * p.nsplit <= 192 && rowgroups * TM < p.MH) rowgroups *= 2;
    p.pitch = ((p.CIN / 4) % 2 == 1) ? p.CIN : p.CIN + 4;          // 4 * odd words
    // weight block: all input channels of a tap when two such blocks fit 44 KB, else fewer
    int kblk = (22 * 1024 / 4 / p.COUT) / 4 * 4;
    if (kblk < 4) return VAE2_ERR_UNSUPPORTED;
    if (kblk > p.CIN) kblk = p.CIN;
    p.kblk = kblk;
    p.nblk = (p.CIN + kblk - 1) / kblk;
    // small images: shrink the patch (never below 128 threads: measured, 64-thread CTAs lose 1.7x) until every
    // SM has a CTA
    auto n_tiles = [&](int pw_, int rg_) {
        return (long long)p.B * ((p.MW + pw_ - 1) / pw_) * ((p.MH + rg_ * TM - 1) / (rg_ * TM));
    };
    while (rowgroups > 1 && pw * (rowgroups / 2) * p.nsplit >= 128 && n_tiles(pw, rowgroups) < kNumSMs) rowgroups /= 2;
    while (pw > 8 && (pw / 2) * rowgroups * p.nsplit >= 128 && n_tiles(pw, rowgroups) < kNumSMs) pw >>= 1;
    // 72-lane layers with too few warps to hide latency (a few fat CTAs, 60 KB patches): the GEMM tiling spreads the
    // same work over more of them (measured 103 vs 154 us at 64x128).
    // The stride-2 data gradient stays here regardless: the GEMM path spends 3/4 of its FMAs on structural zeros.
    if (p.sO == 1 && (p.CIN > 36 || p.COUT > 36) &&
        n_tiles(pw, rowgroups) * ((pw * rowgroups * p.nsplit + 31) / 32) < 8LL * kNumSMs)
        return VAE2_ERR_UNSUPPORTED;
    p.PW = pw;
    size_t smem = 0;
    for (;; rowgroups /= 2) {
        p.PH = rowgroups * TM;
        p.PHin = (p.PH - 1) * p.stride + (dh_max - p.dh_min) + 1;
        p.PWin = (p.PW - 1) * p.stride + (dw_max - p.dw_min) + 1;
        smem = ((size_t)p.PHin * p.PWin * p.pitch + (size_t)2 * kblk * p.COUT) * sizeof(float);
        if (smem <= 72 * 1024 || rowgroups == 1) break;
    }
    if (smem > 200 * 1024) return VAE2_ERR_UNSUPPORTED;
    p.groups = p.PW * rowgroups;
    if (p.groups * p.nsplit > 192) return VAE2_ERR_UNSUPPORTED;
    p.tiles_w = (p.MW + p.PW - 1) / p.PW;
    p.tiles_h = (p.MH + p.PH - 1) / p.PH;
    const int grid = p.B * p.tiles_w * p.tiles_h;
    switch (tn) {
        case 24: return launch_t<24>(p, smem, grid, st);
        case 20: return launch_t<20>(p, smem, grid, st);
        case 16: return launch_t<16>(p, smem, grid, st);
        case 12: return launch_t<12>(p, smem, grid, st);
        case 8: return launch_t<8>(p, smem, grid, st);
        default: return launch_t<4>(p, smem, grid, st);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the narrow layers:  dW[tap][ci][co] += sum_pixels x[pixel*stride + off(tap)][ci] * dy[pixel][co]
//
// The GEMM form (conv_simt.cu, M = taps*Cin rows) pads 180 x 20 outputs to 256 x 32 and re-gathers x per K chunk.
// Here a persistent CTA walks output patches; per patch it stages x (+halo) and dy in shared memory and every
// thread owns a 4(ci) x TN(co) block of one tap's dW in registers for the CTA's whole lifetime:
//   thread = (pixel group, tap, ci quad, co slice);  per pixel: one float4 of x, TN/4 float4 of dy, 4*TN FMAs.
// The pixel groups are folded through shared memory at the end and each CTA issues ONE atomic add per weight
// (fp32 atomics as in the GEMM form: dW must be zero-filled by the caller).
struct WgParams {
    const float* x; const float* dy; float* dw;
    int B, IH, IW, ldx, OH, OW, ldy;
    int CIN, COUT, k, taps, stride, pad;
    int nsplit, rows, groups;          // threads = groups * rows * nsplit, rows = taps * CIN/4
    int PW, PH, lgPW, PWin, PHin, pitchx, pitchd;
    int tiles_w, tiles_h, total_tiles;
};

template <int TN>
__global__ void __launch_bounds__(256, 2)
wgrad_direct_kernel(const WgParams p) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                           // [PHin*PWin][pitchx]
    float* ds = smem + (size_t)p.PHin * p.PWin * p.pitchx;      // [PH*PW][pitchd]
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int per_group = p.rows * p.nsplit;
    const int pg = tid / per_group, rem = tid - pg * per_group;
    const int ns = rem / p.rows, r = rem - ns * p.rows;
    const int q4 = p.CIN / 4;
    const int tap = r / q4, ciq = r - tap * q4;
    const int ky = tap / p.k, kx = tap - ky * p.k;
    const int n0 = ns * TN;
    const int npix = p.PH * p.PW;
    const int per_img = p.tiles_w * p.tiles_h;

    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / per_img;
        const int t = tile - b * per_img;
        const int oh0 = (t / p.tiles_w) * p.PH, ow0 = (t % p.tiles_w) * p.PW;
        __syncthreads();                       // previous patch fully consumed
        {
            const int ih0 = oh0 * p.stride - p.pad, iw0 = ow0 * p.stride - p.pad;
            const int nvec = p.PHin * p.PWin * q4;
            for (int v = tid; v < nvec; v += nthreads) {
                const int pix = v / q4, q = v - pix * q4;
                const int pr = pix / p.PWin, pc = pix - pr * p.PWin;
                const int ih = ih0 + pr, iw = iw0 + pc;
                float* dst = xs + (size_t)pix * p.pitchx + 4 * q;
                if (ih >= 0 && ih < p.IH && iw >= 0 && iw < p.IW)
                    cp_async16(dst, p.x + (((long long)b * p.IH + ih) * p.IW + iw) * p.ldx + 4 * q);
                else
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const int d4 = p.COUT / 4;
            const int nvd = npix * d4;
            for (int v = tid; v < nvd; v += nthreads) {
                const int pix = v / d4, q = v - pix * d4;
                const int pr = pix >> p.lgPW, pc = pix & (p.PW - 1);
                const int oh = oh0 + pr, ow = ow0 + pc;
                float* dst = ds + (size_t)pix * p.pitchd + 4 * q;
                if (oh < p.OH && ow < p.OW)
                    cp_async16(dst, p.dy + (((long long)b * p.OH + oh) * p.OW + ow) * p.ldy + 4 * q);
                else
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);   // zero dy: no contribution
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        if (pg < p.groups) {
            const float* xb = xs + (size_t)(ky * p.PWin + kx) * p.pitchx + 4 * ciq;
            const float* db = ds + n0;
#pragma unroll 1
            for (int px = pg; px < npix; px += p.groups) {
                const int pr = px >> p.lgPW, pc = px & (p.PW - 1);
                const float4 a = *reinterpret_cast<const float4*>(xb + (size_t)(pr * p.stride * p.PWin + pc * p.stride) * p.pitchx);
                float bv[TN];
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 tt = *reinterpret_cast<const float4*>(db + (size_t)px * p.pitchd + j);
                    bv[j] = tt.x; bv[j + 1] = tt.y; bv[j + 2] = tt.z; bv[j + 3] = tt.w;
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    acc[0][j] = fmaf(a.x, bv[j], acc[0][j]);
                    acc[1][j] = fmaf(a.y, bv[j], acc[1][j]);
                    acc[2][j] = fmaf(a.z, bv[j], acc[2][j]);
                    acc[3][j] = fmaf(a.w, bv[j], acc[3][j]);
                }
            }
        }
    }
    // fold the pixel groups (fixed order), then one atomic per weight and CTA
    float* red = smem;                        // [per_group][4*TN]
    for (int g = 1; g < p.groups; ++g) {
        __syncthreads();
        if (pg == g) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) red[(size_t)rem * 4 * TN + i * TN + j] = acc[i][j];
        }
        __syncthreads();
        if (pg == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] += red[(size_t)rem * 4 * TN + i * TN + j];
        }
    }
    if (pg == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* o = p.dw + ((size_t)tap * p.CIN + 4 * ciq + i) * p.COUT + n0;
#pragma unroll
            for (int j = 0; j < TN; ++j) atomicAdd(o + j, acc[i][j]);
        }
    }
}

template <int TN>
static int launch_wg_t(const WgParams& p, size_t smem, int grid, int threads, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(wgrad_direct_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return VAE2_ERR_CUDA;
        attr = true;
    }
    note_kernel("direct::wgrad_direct_kernel");
    wgrad_direct_kernel<TN><<<grid, threads, smem, st>>>(p);
    return check_launch();
}

}  // namespace direct

int conv_fwd_direct(const float* x, const float* wp, const float* bias, float* y, const ConvGeom& g, cudaStream_t st) {
    direct::Params p{};
    p.a = x; p.w = wp; p.bias = bias; p.out = y;
    p.B = g.B; p.IH = g.H; p.IW = g.W; p.lda = g.ldx; p.OH = g.Ho; p.OW = g.Wo; p.ldo = g.ldy;
    p.CIN = g.Cin_p; p.COUT = g.Cout_p; p.taps = g.k * g.k; p.stride = g.stride;
    for (int t = 0; t < p.taps; ++t) { p.dh[t] = t / g.k - g.pad; p.dw[t] = t % g.k - g.pad; p.wt[t] = t; }
    p.MH = g.Ho; p.MW = g.Wo; p.sO = 1; p.oh_off = 0; p.ow_off = 0;
    p.accumulate = 0;
    return direct::launch(p, st);
}

// dx (=|+=) sum_taps dy[(p + pad - tap) / stride] * wpT[tap]   (wpT = [tap][Cout_p][Cin_p]).
// stride 1: one launch with mirrored tap offsets.  stride 2 (3x3): the dx pixels split into the four parity classes
// (h%2, w%2); class (ph, pw) only receives taps with ky = ph+pad, kx = pw+pad (mod 2) and is a small stride-1
// convolution over dy whose results land on the class's strided pixel grid -- no FMA is spent on the structural zeros.
int conv_dgrad_direct(const float* dy, const float* wpT, float* dx, const ConvGeom& g, int accumulate, cudaStream_t st) {
    direct::Params p{};
    p.a = dy; p.w = wpT; p.bias = nullptr; p.out = dx;
    p.B = g.B; p.IH = g.Ho; p.IW = g.Wo; p.lda = g.ldy; p.OH = g.H; p.OW = g.W; p.ldo = g.ldx;
    p.CIN = g.Cout_p; p.COUT = g.Cin_p; p.stride = 1;
    p.accumulate = accumulate;
    if (g.stride == 1) {
        p.taps = g.k * g.k;
        for (int t = 0; t < p.taps; ++t) { p.dh[t] = g.pad - t / g.k; p.dw[t] = g.pad - t % g.k; p.wt[t] = t; }
        p.MH = g.H; p.MW = g.W; p.sO = 1; p.oh_off = 0; p.ow_off = 0;
        return direct::launch(p, st);
    }
    if (g.stride != 2 || g.k != 3) return VAE2_ERR_UNSUPPORTED;
    for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
            direct::Params q = p;
            q.MH = (g.H - ph + 1) / 2; q.MW = (g.W - pw + 1) / 2;
            q.sO = 2; q.oh_off = ph; q.ow_off = pw;
            if (q.MH <= 0 || q.MW <= 0) continue;
            int nt = 0;
            for (int ky = 0; ky < 3; ++ky) {
                if (((ph + g.pad - ky) & 1) != 0) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    if (((pw + g.pad - kx) & 1) != 0) continue;
                    q.dh[nt] = (ph + g.pad - ky) / 2; q.dw[nt] = (pw + g.pad - kx) / 2; q.wt[nt] = ky * 3 + kx;
                    ++nt;
                }
            }
            q.taps = nt;
            const int e = direct::launch(q, st);
            if (e != VAE2_OK) return e;      // (unsupported shapes fail on the first class, before anything is written)
        }
    return VAE2_OK;
}


// dwp (zero-filled by the caller) += weight gradient; VAE2_ERR_UNSUPPORTED -> caller uses the GEMM form
int conv_wgrad_direct(const float* x, const float* dy, float* dwp, const ConvGeom& g, cudaStream_t st) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("VAE2_DIRECT_WGRAD"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled) return VAE2_ERR_UNSUPPORTED;
    if (g.Cin_p > 36 || g.Cout_p > 36 || g.Cin_p % 4 || g.Cout_p % 4) return VAE2_ERR_UNSUPPORTED;
    direct::WgParams p{};
    p.x = x; p.dy = dy; p.dw = dwp;
    p.B = g.B; p.IH = g.H; p.IW = g.W; p.ldx = g.ldx; p.OH = g.Ho; p.OW = g.Wo; p.ldy = g.ldy;
    p.CIN = g.Cin_p; p.COUT = g.Cout_p; p.k = g.k; p.taps = g.k * g.k; p.stride = g.stride; p.pad = g.pad;
    int tn = 0;
    for (int cand : {20, 16, 12, 8, 4})
        if (p.COUT % cand == 0) { tn = cand; break; }
    if (tn == 0) return VAE2_ERR_UNSUPPORTED;
    p.nsplit = p.COUT / tn;
    p.rows = p.taps * (p.CIN / 4);
    const int per_group = p.rows * p.nsplit;
    if (per_group > 256) return VAE2_ERR_UNSUPPORTED;
    p.groups = 256 / per_group;
    const int threads = (p.groups * per_group + 31) / 32 * 32;      // idle tail lanes have pg == groups
    p.PW = 32; p.lgPW = 5;
    while (p.PW > 8 && p.PW / 2 >= p.OW) { p.PW >>= 1; --p.lgPW; }
    p.PH = 8;
    while (p.PH > 1 && p.PH / 2 >= p.OH) p.PH >>= 1;
    p.PHin = (p.PH - 1) * p.stride + p.k;
    p.PWin = (p.PW - 1) * p.stride + p.k;
    p.pitchx = ((p.CIN / 4) % 2 == 1) ? p.CIN : p.CIN + 4;
    p.pitchd = ((p.COUT / 4) % 2 == 1) ? p.COUT : p.COUT + 4;
    size_t smem = ((size_t)p.PHin * p.PWin * p.pitchx + (size_t)p.PH * p.PW * p.pitchd) * sizeof(float);
    const size_t red = (size_t)per_group * 4 * tn * sizeof(float);
    if (red > smem) smem = red;
    if (smem > 100 * 1024) return VAE2_ERR_UNSUPPORTED;
    p.tiles_w = (p.OW + p.PW - 1) / p.PW;
    p.tiles_h = (p.OH + p.PH - 1) / p.PH;
    p.total_tiles = p.B * p.tiles_w * p.tiles_h;
    int grid = 2 * kNumSMs;
    if (grid > p.total_tiles) grid = p.total_tiles;
    switch (tn) {
        case 20: return direct::launch_wg_t<20>(p, smem, grid, threads, st);
        case 16: return direct::launch_wg_t<16>(p, smem, grid, threads, st);
        case 12: return direct::launch_wg_t<12>(p, smem, grid, threads, st);
        case 8: return direct::launch_wg_t<8>(p, smem, grid, threads, st);
        default: return direct::launch_wg_t<4>(p, smem, grid, threads, st);
    }
}

}  // namespace vae2
