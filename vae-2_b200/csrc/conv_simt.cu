// CUDA-core implicit-GEMM convolution: forward, data gradient and weight gradient.
//
// Replaces the cuDNN calls behind the reference's nn.Conv2d sites
// (lib/models/enc_hrnet.py:27-30, 38-41, 70-76, 188-218, 324-337, 381-404, 1010-1017)
// on the exact-fp32 path (fp32 storage, fp32 FMA accumulate, no TF32 truncation), and is the
// on-device cross-check for the tcgen05 path in conv_tc.cu.  All three passes are one tiled
// GEMM  C[M][N] (+)= A[M][K] * Bm[K][N]  whose A operand is gathered on the fly:
//
//   FWD   : M = B*Ho*Wo pixels, N = Cout_p, K = taps*Cin_p   A = x(shifted)   Bm = wp  [tap][Cin_p][Cout_p]
//   DGRAD : M = B*H*W pixels,   N = Cin_p,  K = taps*Cout_p  A = dy(shifted)  Bm = wpT [tap][Cout_p][Cin_p]
//   WGRAD : M = taps*Cin_p,     N = Cout_p, K = B*Ho*Wo      A = x(shifted)^T Bm = dy  (split-K, fp32 atomics)
//
// Tile: BM x BN outputs per 256-thread CTA, BK = 16, 8x4 register tile per thread.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

constexpr int BK = 16;
constexpr int TM = 8, TN = 4;
constexpr int NTHREADS = 256;   // 128x64 and 256x32 tiles; the 128x32 tile runs 128 threads

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    uint2 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
    h[0] = __floats2bfloat162_rn(v.x, v.y);
    h[1] = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = t;
}

template <typename TA, int BM, int BN, int MODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_igemm_kernel(const TA* __restrict__ Asrc, const void* __restrict__ Bsrc, const float* __restrict__ bias,
                  void* __restrict__ Cdst, ConvGeom g, int accumulate, int k_per_split, FastDiv div_wo, FastDiv div_ho) {
    constexpr int NTHREADS = (BM / TM) * (BN / TN);   // shadows the namespace constant: 256, or 128 for the 128x32 tile
    constexpr int AS = BM + 4;               // smem row stride (floats); keeps 16B alignment
    constexpr int A_VECS = BM * BK / 4;      // float4 per A tile
    constexpr int A_PER_T = A_VECS / NTHREADS;
    constexpr int B_VECS = BK * BN / 4;
    constexpr int NTX = BN / TN;             // threads along N
    static_assert(A_VECS % NTHREADS == 0, "A tile / threads");

    __shared__ __align__(16) float As[BK][AS];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int taps = g.k * g.k;

    // GEMM extents
    long long M;
    int N, Cch;  // Cch: channels per tap on the K axis (FWD/DGRAD)
    if (MODE == MODE_FWD)   { M = (long long)g.B * g.Ho * g.Wo; N = g.Cout_p; Cch = g.Cin_p; }
    if (MODE == MODE_DGRAD) { M = (long long)g.B * g.H * g.W;   N = g.Cin_p;  Cch = g.Cout_p; }
    // stride-2 data gradient by parity class (k_per_split carries class+1 in this mode): only the dx pixels with
    // (h%2, w%2) == (ph, pw) are rows of this launch and only the taps that reach them are visited, so no FMA is
    // spent on the structural zeros of the transposed convolution
    const int cls = (MODE == MODE_DGRAD) ? k_per_split - 1 : -1;
    const int ph = cls >= 0 ? (cls >> 1) : 0, pw = cls >= 0 ? (cls & 1) : 0;
    const int OHc = cls >= 0 ? (g.H - ph + 1) / 2 : g.H, OWc = cls >= 0 ? (g.W - pw + 1) / 2 : g.W;
    if (cls >= 0) M = (long long)g.B * OHc * OWc;
    int n_valid = taps;
    unsigned tap_code = 0;            // 4 bits per visited tap
    if (cls >= 0) {
        n_valid = 0;
        for (int ky = 0; ky < g.k; ++ky) {
            if (((ph + g.pad - ky) & 1) != 0) continue;
            for (int kx = 0; kx < g.k; ++kx) {
                if (((pw + g.pad - kx) & 1) != 0) continue;
                tap_code |= (unsigned)(ky * g.k + kx) << (4 * n_valid);
                ++n_valid;
            }
        }
    }
    auto tap_of = [&](int i) { return cls >= 0 ? (int)((tap_code >> (4 * i)) & 15u) : i; };
    if (MODE == MODE_WGRAD) { M = (long long)taps * g.Cin_p;    N = g.Cout_p; Cch = 0; }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // B-tile loader coordinates (one float4 per thread, threads beyond the tile idle)
    const int b_row = tid / (BN / 4), b_col = (tid % (BN / 4)) * 4;
    const bool b_active = tid < B_VECS;

    if (MODE != MODE_WGRAD) {
        // ---- per-thread A rows are fixed pixels: decode once ----
        int a_n[A_PER_T], a_h[A_PER_T], a_w[A_PER_T];
        bool a_ok[A_PER_T];
        const int a_kq = (tid % 4) * 4;  // channel quad inside the BK chunk
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
            const int v = tid + j * NTHREADS;
            const long long m = m0 + v / 4;
            a_ok[j] = m < M;
            const long long mm = a_ok[j] ? m : 0;
            const int OW = (MODE == MODE_FWD) ? g.Wo : OWc;
            const int OH = (MODE == MODE_FWD) ? g.Ho : OHc;
            const int w = (int)(mm % OW);
            const long long t = mm / OW;
            a_h[j] = (int)(t % OH);
            a_n[j] = (int)(t / OH);
            a_w[j] = w;
            if (cls >= 0) { a_h[j] = a_h[j] * 2 + ph; a_w[j] = a_w[j] * 2 + pw; }
        }
        const float* Bw = reinterpret_cast<const float*>(Bsrc);
        const int ldb = N;
        const int IH = (MODE == MODE_FWD) ? g.H : g.Ho;   // spatial extent of the gathered tensor
        const int IW = (MODE == MODE_FWD) ? g.W : g.Wo;
        const int lda = (MODE == MODE_FWD) ? g.ldx : g.ldy;

        // Software pipeline: the global loads of chunk i+1 are issued (into registers) before the FMAs of chunk i,
        // so their latency overlaps the math; shared memory is single-buffered (store after the compute barrier).
        const TA* a_ptr[A_PER_T];
        float4 ar[A_PER_T];
        float4 br;
        auto resolve_tap = [&](int tap) {
            const int ky = tap / g.k, kx = tap - ky * g.k;
#pragma unroll
            for (int j = 0; j < A_PER_T; ++j) {
                int ih, iw;
                bool ok = a_ok[j];
                if (MODE == MODE_FWD) {
                    ih = a_h[j] * g.stride - g.pad + ky;
                    iw = a_w[j] * g.stride - g.pad + kx;
                } else {
                    const int th = a_h[j] + g.pad - ky, tw = a_w[j] + g.pad - kx;
                    ok = ok && th >= 0 && tw >= 0 && (th % g.stride == 0) && (tw % g.stride == 0);
                    ih = th / g.stride;
                    iw = tw / g.stride;
                }
                ok = ok && ih >= 0 && ih < IH && iw >= 0 && iw < IW;
                a_ptr[j] = ok ? Asrc + (((long long)a_n[j] * IH + ih) * IW + iw) * lda : nullptr;
            }
        };
        auto gload = [&](int tap, int kc) {
#pragma unroll
            for (int j = 0; j < A_PER_T; ++j) {
                const int c = kc + a_kq;
                ar[j] = (a_ptr[j] != nullptr && c < Cch) ? load4<TA>(a_ptr[j] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            br = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_active && kc + b_row < Cch && n0 + b_col < N)
                br = *reinterpret_cast<const float4*>(Bw + ((long long)tap * Cch + kc + b_row) * ldb + n0 + b_col);
        };
        auto sstore = [&]() {
#pragma unroll
            for (int j = 0; j < A_PER_T; ++j) {
                const int r = (tid + j * NTHREADS) / 4;
                As[a_kq + 0][r] = ar[j].x; As[a_kq + 1][r] = ar[j].y;
                As[a_kq + 2][r] = ar[j].z; As[a_kq + 3][r] = ar[j].w;
            }
            if (b_active) *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = br;
        };
        const int nkc = (Cch + BK - 1) / BK;
        const int n_iter = n_valid * nkc;
        int tap = 0, kc = 0;             // tap = position in the visited-tap sequence
        resolve_tap(tap_of(0));
        gload(tap_of(0), 0);
        sstore();
        __syncthreads();
        for (int it = 0; it < n_iter; ++it) {
            int ntap = tap, nkcv = kc + BK;
            if (nkcv >= Cch) { nkcv = 0; ntap = tap + 1; }
            const bool has_next = it + 1 < n_iter;
            if (has_next) {
                if (ntap != tap) resolve_tap(tap_of(ntap));
                gload(tap_of(ntap), nkcv);
            }
            // only the channels that exist in this chunk (Cch is a multiple of 4): an 18->20-lane
            // tensor costs 20 k-steps per tap, not 32
            const int kmax = (Cch - kc) < BK ? (Cch - kc) : BK;
            for (int k4 = 0; k4 < kmax; k4 += 4) {
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    const int kk = k4 + kq;
                    const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
                    const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
            }
            __syncthreads();
            if (has_next) {
                sstore();
                __syncthreads();
            }
            tap = ntap; kc = nkcv;
        }
        // ---- epilogue: rows are output pixels ----
        TA* out = reinterpret_cast<TA*>(Cdst);
        const int ldo = (MODE == MODE_FWD) ? g.ldy : g.ldx;
        const int n = n0 + tx * TN;
        if (n < N) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == MODE_FWD && bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const long long m = m0 + ty * TM + i;
                if (m < M) {
                    float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
                    long long pix = m;
                    if (cls >= 0) {                 // class row -> dx pixel
                        const int ww = (int)(m % OWc);
                        const long long t = m / OWc;
                        const int hh = (int)(t % OHc);
                        pix = ((t / OHc) * g.H + hh * 2 + ph) * g.W + ww * 2 + pw;
                    }
                    TA* p = out + pix * ldo + n;
                    if (MODE == MODE_DGRAD && accumulate) {
                        const float4 o = load4<TA>(p);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    store4<TA>(p, v);
                }
            }
        }
    } else {
        // ---- WGRAD: A rows are (tap, ci) quads; K runs over output pixels of this split ----
        const long long Ktot = (long long)g.B * g.Ho * g.Wo;
        const long long k_begin = (long long)blockIdx.z * k_per_split;
        long long k_end = k_begin + k_per_split;
        if (k_end > Ktot) k_end = Ktot;
        const TA* Bd = reinterpret_cast<const TA*>(Bsrc);  // dy
        constexpr int MQ = BM / 4;
        int a_tap_ky[A_PER_T], a_tap_kx[A_PER_T], a_ci[A_PER_T], a_kk[A_PER_T], a_mq[A_PER_T];
        bool a_ok[A_PER_T];
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
            const int v = tid + j * NTHREADS;
            a_mq[j] = v % MQ;
            a_kk[j] = v / MQ;
            const long long m = m0 + a_mq[j] * 4;
            a_ok[j] = m < M;
            const int mm = a_ok[j] ? (int)m : 0;
            const int tap = mm / g.Cin_p;
            a_ci[j] = mm - tap * g.Cin_p;
            a_tap_ky[j] = tap / g.k;
            a_tap_kx[j] = tap - a_tap_ky[j] * g.k;
        }
        float4 ar[A_PER_T];
        float4 br;
        auto gload = [&](long long k0) {
#pragma unroll
            for (int j = 0; j < A_PER_T; ++j) {
                const long long p = k0 + a_kk[j];
                ar[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_ok[j] && p < k_end) {
                    if (g.k == 1 && g.stride == 1) {
                        ar[j] = load4<TA>(Asrc + p * g.ldx + a_ci[j]);          // 1x1: output pixel == input pixel
                    } else {
                        const unsigned t = fast_div((unsigned)p, div_wo);       // pixel index < 2^31 (checked by the launcher)
                        const int wo = (int)((unsigned)p - t * div_wo.d);
                        const unsigned nb = fast_div(t, div_ho);
                        const int ho = (int)(t - nb * div_ho.d);
                        const int ih = ho * g.stride - g.pad + a_tap_ky[j];
                        const int iw = wo * g.stride - g.pad + a_tap_kx[j];
                        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                            ar[j] = load4<TA>(Asrc + (((long long)nb * g.H + ih) * g.W + iw) * g.ldx + a_ci[j]);
                    }
                }
            }
            br = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_active && k0 + b_row < k_end && n0 + b_col < N)
                br = load4<TA>(Bd + (k0 + b_row) * g.ldy + n0 + b_col);
        };
        auto sstore = [&]() {
#pragma unroll
            for (int j = 0; j < A_PER_T; ++j)
                *reinterpret_cast<float4*>(&As[a_kk[j]][a_mq[j] * 4]) = ar[j];
            if (b_active) *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = br;
        };
        if (k_begin < k_end) {
            gload(k_begin);
            sstore();
        }
        __syncthreads();
        for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool has_next = k0 + BK < k_end;
            if (has_next) gload(k0 + BK);       // in flight during the FMAs below
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
            if (has_next) {
                sstore();
                __syncthreads();
            }
        }
        float* out = reinterpret_cast<float*>(Cdst);  // dwp [M][N]
        const int n = n0 + tx * TN;
        if (n < N) {
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const long long m = m0 + ty * TM + i;
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < TN; ++j) atomicAdd(out + m * N + n + j, acc[i][j]);
                }
            }
        }
    }
}

// Pick the N tile that wastes the fewest padded columns (ties -> wider tile).
static inline int pick_bn(int N) {
    const int w64 = ((N + 63) / 64) * 64, w32 = ((N + 31) / 32) * 32;
    return (w32 < w64) ? 32 : 64;
}

static int check_geom(const ConvGeom& g, int dtype) {
    const int a = (dtype == VAE2_DT_F32) ? 4 : 8;
    if (g.Cin_p % a || g.Cout_p % a || g.ldx % a || g.ldy % a) return VAE2_ERR_ARG;
    if (g.k != 1 && g.k != 3) return VAE2_ERR_UNSUPPORTED;
    if (g.stride != 1 && g.stride != 2) return VAE2_ERR_UNSUPPORTED;
    return VAE2_OK;
}

template <typename TA, int MODE>
static int launch_igemm(const void* A, const void* Bm, const float* bias, void* C, const ConvGeom& g, int accumulate,
                        cudaStream_t st, int cls = -1) {
    const int taps = g.k * g.k;
    long long M;
    int N;
    if (MODE == MODE_FWD) { M = (long long)g.B * g.Ho * g.Wo; N = g.Cout_p; }
    else if (MODE == MODE_DGRAD) { M = (long long)g.B * g.H * g.W; N = g.Cin_p; }
    else { M = (long long)taps * g.Cin_p; N = g.Cout_p; }
    if (MODE == MODE_DGRAD && cls >= 0) M = (long long)g.B * ((g.H - (cls >> 1) + 1) / 2) * ((g.W - (cls & 1) + 1) / 2);
    if (M <= 0) return VAE2_OK;
    const int kcls = cls + 1;
    const FastDiv div_wo = make_fastdiv((unsigned)g.Wo), div_ho = make_fastdiv((unsigned)g.Ho);
    if ((long long)g.B * g.Ho * g.Wo >= (1LL << 31)) return VAE2_ERR_UNSUPPORTED;
    int bn = pick_bn(N);
    if (MODE == MODE_WGRAD) {
        // both tile shapes pad M (taps*Cin_p rows) and N; take the one that wastes less work
        const long long w64 = ((M + 127) / 128 * 128) * ((N + 63) / 64 * 64), w32 = ((M + 255) / 256 * 256) * ((N + 31) / 32 * 32);
        bn = (w64 <= w32) ? 64 : 32;
    }
    if (MODE != MODE_WGRAD) {
        if (bn == 64) {
            dim3 grid((unsigned)((M + 127) / 128), (N + 63) / 64);
            note_kernel("conv_igemm_kernel<128,64>");
            conv_igemm_kernel<TA, 128, 64, MODE><<<grid, NTHREADS, 0, st>>>((const TA*)A, Bm, bias, C, g, accumulate, kcls, div_wo, div_ho);
        } else if (((M + 255) / 256) * ((N + 31) / 32) < 2 * kNumSMs) {
            // small layer: half-height tiles (128 threads) so that the grid still covers the GPU about twice
            dim3 grid((unsigned)((M + 127) / 128), (N + 31) / 32);
            note_kernel("conv_igemm_kernel<128,32>");
            conv_igemm_kernel<TA, 128, 32, MODE><<<grid, 128, 0, st>>>((const TA*)A, Bm, bias, C, g, accumulate, kcls, div_wo, div_ho);
        } else {
            dim3 grid((unsigned)((M + 255) / 256), (N + 31) / 32);
            note_kernel("conv_igemm_kernel<256,32>");
            conv_igemm_kernel<TA, 256, 32, MODE><<<grid, NTHREADS, 0, st>>>((const TA*)A, Bm, bias, C, g, accumulate, kcls, div_wo, div_ho);
        }
    } else {
        const long long Ktot = (long long)g.B * g.Ho * g.Wo;
        const int bm = (bn == 64) ? 128 : 256;
        const long long tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
        long long splits = (4LL * kNumSMs + tiles - 1) / tiles;           // ~4 CTAs per SM in flight
        const long long max_splits = (Ktot + 4 * BK - 1) / (4 * BK);      // at least 4 K-chunks per split
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        long long kps = (Ktot + splits - 1) / splits;
        kps = ((kps + BK - 1) / BK) * BK;
        splits = (Ktot + kps - 1) / kps;
        dim3 grid((unsigned)((M + bm - 1) / bm), (N + bn - 1) / bn, (unsigned)splits);
        note_kernel(bn == 64 ? "conv_igemm_kernel<128,64,wgrad>" : "conv_igemm_kernel<256,32,wgrad>");
        if (bn == 64)
            conv_igemm_kernel<TA, 128, 64, MODE><<<grid, NTHREADS, 0, st>>>((const TA*)A, Bm, nullptr, C, g, 0, (int)kps, div_wo, div_ho);
        else
            conv_igemm_kernel<TA, 256, 32, MODE><<<grid, NTHREADS, 0, st>>>((const TA*)A, Bm, nullptr, C, g, 0, (int)kps, div_wo, div_ho);
    }
    return check_launch();
}

int conv_fwd_simt(const void* x, const float* wp, const float* bias, void* y, int dtype, const ConvGeom& g, cudaStream_t st) {
    if (int e = check_geom(g, dtype)) return e;
    if (dtype == VAE2_DT_F32) {
        const int e = conv_fwd_direct((const float*)x, wp, bias, (float*)y, g, st);   // register-tiled direct conv first
        if (e != VAE2_ERR_UNSUPPORTED) return e;
        return launch_igemm<float, MODE_FWD>(x, wp, bias, y, g, 0, st);
    }
    return launch_igemm<__nv_bfloat16, MODE_FWD>(x, wp, bias, y, g, 0, st);
}

int conv_dgrad_simt(const void* dy, const float* wpT, void* dx, int dtype, const ConvGeom& g, int accumulate, cudaStream_t st) {
    if (int e = check_geom(g, dtype)) return e;
    if (dtype == VAE2_DT_F32) {
        const int e = conv_dgrad_direct((const float*)dy, wpT, (float*)dx, g, accumulate, st);
        if (e != VAE2_ERR_UNSUPPORTED) return e;
    }
    if (g.stride == 2 && g.k == 3) {            // four parity-class launches (see the kernel)
        for (int cls = 0; cls < 4; ++cls) {
            const int e = dtype == VAE2_DT_F32 ? launch_igemm<float, MODE_DGRAD>(dy, wpT, nullptr, dx, g, accumulate, st, cls)
                                               : launch_igemm<__nv_bfloat16, MODE_DGRAD>(dy, wpT, nullptr, dx, g, accumulate, st, cls);
            if (e != VAE2_OK) return e;
        }
        return VAE2_OK;
    }
    if (dtype == VAE2_DT_F32) return launch_igemm<float, MODE_DGRAD>(dy, wpT, nullptr, dx, g, accumulate, st);
    return launch_igemm<__nv_bfloat16, MODE_DGRAD>(dy, wpT, nullptr, dx, g, accumulate, st);
}

// dwp must be zero-filled by the caller (split-K partial sums are added atomically).
int conv_wgrad_simt(const void* x, const void* dy, float* dwp, int dtype, const ConvGeom& g, cudaStream_t st) {
    if (int e = check_geom(g, dtype)) return e;
    if (dtype == VAE2_DT_F32) {
        const int e = conv_wgrad_direct((const float*)x, (const float*)dy, dwp, g, st);
        if (e != VAE2_ERR_UNSUPPORTED) return e;
        return launch_igemm<float, MODE_WGRAD>(x, dy, nullptr, dwp, g, 0, st);
    }
    return launch_igemm<__nv_bfloat16, MODE_WGRAD>(x, dy, nullptr, dwp, g, 0, st);
}

// dbias[c] (=|+=) sum over pixels of dy[p][c]  (heads' biased 1x1 convs, enc_hrnet.py:324-337)
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, float* __restrict__ dbias, long long P, int C, int ld) {
    // block handles a pixel range for all channels; threads = 32 channel lanes x 8 pixel rows
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float s = 0.f;
        if (c < C)
            for (long long p = blockIdx.x * 8LL + row; p < P; p += (long long)gridDim.x * 8) s += to_f<T>(dy[p * ld + c]);
        red[row][lane] = s;
        __syncthreads();
        if (row == 0 && c < C) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) t += red[r][lane];
            atomicAdd(dbias + c, t);
        }
        __syncthreads();
    }
}

// Vectorised variant: a thread owns one 16-byte piece of the lane range and walks the block's pixel range with it
// (the scalar kernel above re-walked the pixels once per 32 channels with 2-byte loads: 885 GB/s on the 270-lane heads).
template <typename T, int V>
__global__ void __launch_bounds__(256)
bias_grad_vec_kernel(const T* __restrict__ dy, float* __restrict__ dbias, long long P, int C, int ld, int nc, long long per_block) {
    __shared__ float red[2048];
    const int R = 256 / nc;                               // pixel rows in flight per block
    const int chunk = threadIdx.x % nc, r = threadIdx.x / nc;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const long long p0 = blockIdx.x * per_block, p1 = min(P, p0 + per_block);
    if (r < R) {
        for (long long p = p0 + r; p < p1; p += R) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(dy + p * ld + chunk * V));
            if constexpr (V == 4) {
                acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y);
                acc[2] += __uint_as_float(q.z); acc[3] += __uint_as_float(q.w);
            } else {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
            }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) red[(r * nc + chunk) * V + e] = acc[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
        for (int rr = 0; rr < R; ++rr) t += red[rr * nc * V + c];
        atomicAdd(dbias + c, t);
    }
}

int bias_grad(const void* dy, float* dbias, int dtype, long long P, int C, int ld, int accumulate, cudaStream_t st) {
    if (!accumulate) {
        if (cudaMemsetAsync(dbias, 0, sizeof(float) * C, st) != cudaSuccess) return VAE2_ERR_CUDA;
    }
    const int V = dtype == VAE2_DT_F32 ? 4 : 8;
    const int nc = (C + V - 1) / V;                       // 16-byte pieces that hold real channels (the tail piece may read pad lanes)
    if (ld % V == 0 && nc * V <= ld && nc <= 256 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
        int grid = 4 * kNumSMs;
        if ((long long)grid * 64 > P) grid = (int)((P + 63) / 64);
        const long long per_block = (P + grid - 1) / grid;
        if (dtype == VAE2_DT_F32)
            bias_grad_vec_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)dy, dbias, P, C, ld, nc, per_block);
        else
            bias_grad_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, dbias, P, C, ld, nc, per_block);
        return check_launch();
    }
    const int grid = stream_grid(P, 8 * 64, 2);
    if (dtype == VAE2_DT_F32)
        bias_grad_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, dbias, P, C, ld);
    else
        bias_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, dbias, P, C, ld);
    return check_launch();
}

}  // namespace vae2
