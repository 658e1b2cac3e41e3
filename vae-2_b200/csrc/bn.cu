// Batch normalisation for channels-last tensors: Welford statistics, finalize (+ running
// stats, momentum 0.01 / eps 1e-5 as at every BN site of lib/models/enc_hrnet.py:22-23),
// fused apply (+residual)(+ReLU), and the two-pass backward.
//
// Semantics follow torch.nn.BatchNorm2d / SyncBatchNorm as the reference uses them
// (tools/train.py:217-218): training mode normalises with the biased batch variance and
// updates running_var with the unbiased one; under SyncBN the per-rank (count, mean, M2)
// partials are all-gathered by the host (NCCL) and merged with the same Chan formula that
// merges the per-CTA partials here.
//
// Threading: a pixel row of Cp channels is covered by `lanes = Cp / V` threads, each owning V
// consecutive channels (16-byte vectors); a 256-thread CTA covers R = 256 / lanes pixels per
// step and strides over the image, so every global access is a full 16-byte vector and
// consecutive threads touch consecutive addresses.
//
// Single-rank training BNs run as ONE cooperative launch per direction (bn_fwd_fused / bn_bwd_fused):
// statistics -> grid barrier -> per-channel finalize spread over the grid's warps -> grid barrier ->
// elementwise pass.  The stand-alone kernels stay for SyncBN (a collective sits between the phases)
// and for eval-mode statistics.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace vae2 {

constexpr int BN_THREADS = 256;
constexpr int BN_SPLIT_PARTS = kNumSMs * 2;  // stand-alone reductions: few, fat partials keep the merge kernels short
constexpr int BN_MAX_PARTS = kNumSMs * 4;    // fused kernels: up to 4 co-resident CTAs per SM

// Thread layout shared by every pass: `lanes` threads cover one pixel's Cp channels (V each), the CTA
// covers R pixels per step; a thread keeps its channel group for the whole kernel, so per-channel
// coefficients live in registers.
template <typename T>
struct Lay {
    static constexpr int V = Vec<T>::N;
    int lanes, R, lane, r;
    int bid, nb;      // this CTA's index / the CTA count WITHIN ITS STATISTICS GROUP (the whole grid when there is one group)
    bool active;
    __device__ __forceinline__ explicit Lay(int Cp, int bid_ = -1, int nb_ = 0) {
        lanes = Cp / V;
        R = blockDim.x / lanes;
        lane = threadIdx.x % lanes;
        r = threadIdx.x / lanes;
        active = r < R;
        bid = bid_ < 0 ? (int)blockIdx.x : bid_;
        nb = bid_ < 0 ? (int)gridDim.x : nb_;
    }
};

int bn_stats_max_partials() { return BN_MAX_PARTS; }

// ---------------------------------------------------------------------------
// statistics: per-CTA (count, mean, M2) per channel  -> partials[cta][3][Cp]
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Streaming loop shared by every pass.  A thread visits pixels first, first+stride, ...; its 16-byte vectors
// are fetched with cp.async into the thread's OWN shared-memory slots D iterations ahead and read back right
// before use.  The loads in flight therefore live in shared memory (kPipeBytes per CTA), not in registers:
// ~100 KB per SM in flight at 3 CTAs/SM, which is what HBM needs, while the kernels stay under 85 registers.
// No block-level synchronisation is involved (a thread only ever reads slots it filled itself).
// `reverse` walks the list backwards: the second pass of a fused kernel then starts with the lines the first
// pass touched last, i.e. the ones still in L2.
// ---------------------------------------------------------------------------
constexpr int kPipeSlots = 12;                                  // 16-byte slots per thread
constexpr int kPipeBytes = kPipeSlots * BN_THREADS * 16;        // 48 KB

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int NTENS, typename Body>
__device__ __forceinline__ void stream_pixels(float* sm, const Lay<T>& L, long long P, bool reverse,
                                              const T* const (&base)[NTENS], const int (&ld)[NTENS], Body body) {
    constexpr int V = Vec<T>::N;
    constexpr int D = kPipeSlots / NTENS;
    if (!L.active) return;
    const long long stride = (long long)L.nb * L.R;
    const long long first = (long long)L.bid * L.R + L.r;
    const long long n = first < P ? (P - first + stride - 1) / stride : 0;
    const int c0 = L.lane * V;
    uint4* mine = reinterpret_cast<uint4*>(sm) + threadIdx.x;     // slot s, tensor t: mine[(s*NTENS + t) * BN_THREADS]
    int s_in = 0;
    for (int i = 0; i < D - 1; ++i) {
        if (i < n) {
            const long long q = first + (reverse ? n - 1 - i : (long long)i) * stride;
#pragma unroll
            for (int t = 0; t < NTENS; ++t) cp_async16(mine + (s_in * NTENS + t) * BN_THREADS, base[t] + q * ld[t] + c0);
        }
        cp_async_commit();
        s_in = s_in + 1 == D ? 0 : s_in + 1;
    }
    int s_out = 0;
    for (long long i = 0; i < n; ++i) {
        const long long j = i + D - 1;
        if (j < n) {
            const long long q = first + (reverse ? n - 1 - j : j) * stride;
#pragma unroll
            for (int t = 0; t < NTENS; ++t) cp_async16(mine + (s_in * NTENS + t) * BN_THREADS, base[t] + q * ld[t] + c0);
        }
        cp_async_commit();
        s_in = s_in + 1 == D ? 0 : s_in + 1;
        cp_async_wait<D - 1>();
        uint4 v[NTENS];
#pragma unroll
        for (int t = 0; t < NTENS; ++t) v[t] = mine[(s_out * NTENS + t) * BN_THREADS];
        s_out = s_out + 1 == D ? 0 : s_out + 1;
        body(first + (reverse ? n - 1 - i : i) * stride, v);
    }
    cp_async_wait<0>();
}

template <typename T>
__device__ __forceinline__ void stats_pass(const T* __restrict__ y, long long P, int ld, const Lay<T>& L, float* sm,
                                           float (&mean)[Vec<T>::N], float (&m2)[Vec<T>::N], float& n) {
    constexpr int V = Vec<T>::N;
#pragma unroll
    for (int i = 0; i < V; ++i) { mean[i] = 0.f; m2[i] = 0.f; }
    n = 0.f;
    const T* const base[1] = {y};
    const int lds[1] = {ld};
    stream_pixels<T, 1>(sm, L, P, false, base, lds, [&](long long, const uint4 (&v)[1]) {
        n += 1.f;
        const float inv = 1.f / n;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float xv = elem<T>(v[0], i);
            const float d = xv - mean[i];
            mean[i] += d * inv;
            m2[i] = fmaf(d, xv - mean[i], m2[i]);
        }
    });
}

// CTA merge (Chan) over the thread rows r = 0..R-1 of each lane: a log2(R)-step tree in shared memory
// (sm: [blockDim][2V+1]); the result is valid in the r == 0 threads.
template <int V>
__device__ __forceinline__ void cta_merge_welford(float* sm, int lanes, int R, int lane, int r, bool active,
                                                  float (&mean)[V], float (&m2)[V], float& n) {
    __syncthreads();      // the scratch aliases the streaming slots of threads that may still be reading theirs
    float* mine = sm + threadIdx.x * (2 * V + 1);
#pragma unroll
    for (int i = 0; i < V; ++i) { mine[i] = mean[i]; mine[V + i] = m2[i]; }
    mine[2 * V] = n;
    int h = 1;
    while (h < R) h <<= 1;
    int rows = R;
    for (h >>= 1; h >= 1; h >>= 1) {
        __syncthreads();
        if (active && r < h && r + h < rows) {
            const float* o = sm + ((r + h) * lanes + lane) * (2 * V + 1);
            const float nb = o[2 * V];
            if (nb > 0.f) {
                const float nn = n + nb;
                const float f = nb / nn;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float d = o[i] - mean[i];
                    mean[i] += d * f;
                    m2[i] += o[V + i] + d * d * n * f;
                }
                n = nn;
#pragma unroll
                for (int i = 0; i < V; ++i) { mine[i] = mean[i]; mine[V + i] = m2[i]; }
                mine[2 * V] = n;
            }
        }
        rows = rows < h ? rows : h;
    }
}

template <typename T>
__device__ __forceinline__ void stats_to_partials(const T* __restrict__ y, float* __restrict__ partials, long long P,
                                                  int Cp, int ld, float* sm, int bid = -1, int nb = 0) {
    constexpr int V = Vec<T>::N;
    const Lay<T> L(Cp, bid, nb);
    float mean[V], m2[V], n;
    stats_pass<T>(y, P, ld, L, sm, mean, m2, n);
    cta_merge_welford<V>(sm, L.lanes, L.R, L.lane, L.r, L.active, mean, m2, n);
    if (L.active && L.r == 0) {
        float* out = partials + (long long)blockIdx.x * 3 * Cp;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = L.lane * V + i;
            out[c] = n;
            out[Cp + c] = mean[i];
            out[2 * Cp + c] = m2[i];
        }
    }
}

// ---------------------------------------------------------------------------
// statistics: per-CTA (count, mean, M2) per channel  -> partials[cta][3][Cp]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const T* __restrict__ y, float* __restrict__ partials, long long P, int Cp, int ld) {
    extern __shared__ __align__(16) float sm[];          // [BN_THREADS][2*V + 1]
    stats_to_partials<T>(y, partials, P, Cp, ld, sm);
}

__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
    if (nb <= 0.f) return;
    const float nn = n + nb;
    const float f = nb / nn;
    const float d = mb - mean;
    mean += d * f;
    m2 += m2b + d * d * n * f;
    n = nn;
}

// 5-step butterfly of Chan merges; the symmetric form leaves every lane with the same result
__device__ __forceinline__ void warp_tree_merge(float& n, float& mean, float& m2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o);
        const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
        const float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        const float nn = n + nb;
        if (nn > 0.f) {
            const float d = mb - mean;
            const float f = nb / nn;
            mean = (nb > 0.f && n > 0.f) ? mean + d * f : (nb > 0.f ? mb : mean);
            m2 = m2 + m2b + ((nb > 0.f && n > 0.f) ? d * d * n * f : 0.f);
        }
        n = nn;
    }
}

// One WARP per channel: lanes fold partials lane, lane+32, ... then a 5-step shuffle tree of Chan
// merges (the merge is associative), so the latency is ~n_parts/32 dependent loads, not n_parts.
__device__ __forceinline__ void warp_merge_parts(const float* __restrict__ parts, int n_parts, int Cp, int c,
                                                 float& n, float& mean, float& m2, long long part_stride = 0) {
    if (part_stride == 0) part_stride = 3LL * Cp;
    const int lane = threadIdx.x & 31;
    n = 0.f; mean = 0.f; m2 = 0.f;
    for (int k0 = lane; k0 < n_parts; k0 += 256) {   // 8 independent partials in flight per step
        float a[8][3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + 32 * j;
            const bool valid = k < n_parts;
            const float* p = parts + (long long)(valid ? k : k0) * part_stride;
            // __ldcg: in the fused kernels these were written by other CTAs earlier in the same launch
            a[j][0] = valid ? __ldcg(p + c) : 0.f;          // count 0: skipped by the merge
            a[j][1] = __ldcg(p + Cp + c);
            a[j][2] = __ldcg(p + 2 * Cp + c);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) chan_merge(n, mean, m2, a[j][0], a[j][1], a[j][2]);
    }
    warp_tree_merge(n, mean, m2);
}

// merge n_parts partial sets -> one (count, mean, M2) set  (per-rank partial for SyncBN all-gather)
__global__ void bn_merge_kernel(const float* __restrict__ parts, int n_parts, int Cp, float* __restrict__ merged) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= Cp) return;
    float n, mean, m2;
    warp_merge_parts(parts, n_parts, Cp, c, n, mean, m2);
    if ((threadIdx.x & 31) == 0) { merged[c] = n; merged[Cp + c] = mean; merged[2 * Cp + c] = m2; }
}

// merge + produce normalisation coefficients + running-stat update for channel c (called by a full warp)
__device__ __forceinline__ void finalize_channel(int c, const float* __restrict__ parts, int n_parts, int C, int Cp,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float* __restrict__ running_mean, float* __restrict__ running_var,
                                                 float momentum, float eps, float* __restrict__ mean_o,
                                                 float* __restrict__ invstd_o, float* __restrict__ scale_o,
                                                 float* __restrict__ shift_o, long long part_stride) {
    const bool lead = (threadIdx.x & 31) == 0;
    if (c >= C) {
        if (lead) { mean_o[c] = 0.f; invstd_o[c] = 0.f; scale_o[c] = 0.f; shift_o[c] = 0.f; }
        return;
    }
    // parameters first: their loads overlap the partial merge instead of trailing it
    const float gm = gamma[c], bt = beta[c];
    const float rm0 = running_mean != nullptr ? running_mean[c] : 0.f;
    const float rv0 = running_mean != nullptr ? running_var[c] : 0.f;
    float n, mean, m2;
    warp_merge_parts(parts, n_parts, Cp, c, n, mean, m2, part_stride);
    if (!lead) return;
    const float var = m2 / n;
    const float invstd = rsqrtf(var + eps);
    const float sc = gm * invstd;
    mean_o[c] = mean; invstd_o[c] = invstd; scale_o[c] = sc; shift_o[c] = bt - mean * sc;
    if (running_mean != nullptr) {
        const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
        running_mean[c] = (1.f - momentum) * rm0 + momentum * mean;
        running_var[c] = (1.f - momentum) * rv0 + momentum * unbiased;
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ parts, int n_parts, int C, int Cp,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float momentum, float eps,
                                   float* __restrict__ mean_o, float* __restrict__ invstd_o,
                                   float* __restrict__ scale_o, float* __restrict__ shift_o, long long part_stride) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = gt >> 5;
    if (gt == 0 && nbt != nullptr) *nbt += 1;
    if (c >= Cp) return;
    finalize_channel(c, parts, n_parts, C, Cp, gamma, beta, running_mean, running_var, momentum, eps, mean_o, invstd_o,
                     scale_o, shift_o, part_stride);
}

__global__ void bn_eval_coeffs_kernel(int C, int Cp, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale_o, float* __restrict__ shift_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c >= C) { scale_o[c] = 0.f; shift_o[c] = 0.f; return; }
    const float sc = gamma[c] * rsqrtf(rv[c] + eps);
    scale_o[c] = sc; shift_o[c] = beta[c] - rm[c] * sc;
}

// ---------------------------------------------------------------------------
// apply: out = [relu]( scale*y + shift [+ res] )
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void apply_pass(const T* __restrict__ y, const T* __restrict__ res, T* __restrict__ out,
                                           long long P, int ld_y, int ld_res, int ld_out,
                                           const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                           const Lay<T>& L, float* sm, bool reverse) {
    constexpr int V = Vec<T>::N;
    if (!L.active) return;
    float sc[V], sh[V];
#pragma unroll
    // (fused kernels: written by another CTA before the grid barrier and first read here, so the L1 path is safe;
    //  going to L2 for these few hot addresses from every thread of the grid serialises on one L2 slice)
    for (int k = 0; k < V; ++k) { sc[k] = scale[L.lane * V + k]; sh[k] = shift[L.lane * V + k]; }
    const int c0 = L.lane * V;
    if (res != nullptr) {
        const T* const base[2] = {y, res};
        const int lds[2] = {ld_y, ld_res};
        stream_pixels<T, 2>(sm, L, P, reverse, base, lds, [&](long long q, const uint4 (&v)[2]) {
            Vec<T> o;
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] = fmaf(elem<T>(v[0], k), sc[k], sh[k]) + elem<T>(v[1], k);
            if (relu) {
#pragma unroll
                for (int k = 0; k < V; ++k) o.v[k] = fmaxf(o.v[k], 0.f);
            }
            o.store(out + q * ld_out + c0);
        });
    } else {
        const T* const base[1] = {y};
        const int lds[1] = {ld_y};
        stream_pixels<T, 1>(sm, L, P, reverse, base, lds, [&](long long q, const uint4 (&v)[1]) {
            Vec<T> o;
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] = fmaf(elem<T>(v[0], k), sc[k], sh[k]);
            if (relu) {
#pragma unroll
                for (int k = 0; k < V; ++k) o.v[k] = fmaxf(o.v[k], 0.f);
            }
            o.store(out + q * ld_out + c0);
        });
    }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const T* __restrict__ y, const T* __restrict__ res, T* __restrict__ out, long long P, int Cp,
                int ld_y, int ld_res, int ld_out, const float* __restrict__ scale, const float* __restrict__ shift,
                int relu) {
    extern __shared__ __align__(16) float sm[];
    const Lay<T> L(Cp);
    apply_pass<T>(y, res, out, P, ld_y, ld_res, ld_out, scale, shift, relu, L, sm, false);
}

// ---------------------------------------------------------------------------
// backward pass 1: per-CTA  sum(dyb), sum(dyb * xhat)   dyb = g * [a > 0]
// partials[cta][2][Cp]
// ---------------------------------------------------------------------------
template <typename T, int RELU>
__device__ __forceinline__ void bwd_reduce_to_partials_t(const T* __restrict__ g, const T* __restrict__ a,
                                                       const T* __restrict__ y, float* __restrict__ partials,
                                                       long long P, int Cp, int ld_g, int ld_a, int ld_y,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       float* sm, const float* __restrict__ scale = nullptr,
                                                       const float* __restrict__ shift = nullptr, int bid = -1, int nb = 0) {
    constexpr int relu = RELU;
    // relu: 0 none | 1 mask = [a > 0] read from the stored activation | 2 mask recomputed as
    // [fma(y, scale, shift) > 0] -- the very expression the forward pass rounded into `a` (no residual), so the
    // activation is not read at all
    constexpr int V = Vec<T>::N;
    const Lay<T> L(Cp, bid, nb);
    // s2 accumulates sum(dyb * (y - mu)); the factor invstd is applied once at the end.  Centring on mu keeps
    // the sum free of the cancellation a raw sum(dyb * y) would have.
    float s1[V], s2[V], mu[V], sc[V], sh[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    if (L.active) {
#pragma unroll
        for (int i = 0; i < V; ++i) mu[i] = mean[L.lane * V + i];
        if (relu == 2) {
#pragma unroll
            for (int i = 0; i < V; ++i) { sc[i] = scale[L.lane * V + i]; sh[i] = shift[L.lane * V + i]; }
        }
        auto acc = [&](const uint4& gq, const uint4& yq, const uint4& aq) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float yy = elem<T>(yq, i);
                bool on = true;
                if (relu == 1) on = elem<T>(aq, i) > 0.f;
                if (relu == 2) on = fmaf(yy, sc[i], sh[i]) > 0.f;
                const float gg = on ? elem<T>(gq, i) : 0.f;
                s1[i] += gg;
                s2[i] = fmaf(gg, yy - mu[i], s2[i]);
            }
        };
        if (relu == 1) {
            const T* const base[3] = {g, y, a};
            const int lds[3] = {ld_g, ld_y, ld_a};
            stream_pixels<T, 3>(sm, L, P, false, base, lds, [&](long long, const uint4 (&v)[3]) { acc(v[0], v[1], v[2]); });
        } else {
            const T* const base[2] = {g, y};
            const int lds[2] = {ld_g, ld_y};
            stream_pixels<T, 2>(sm, L, P, false, base, lds, [&](long long, const uint4 (&v)[2]) { acc(v[0], v[1], v[1]); });
        }
#pragma unroll
        for (int i = 0; i < V; ++i) s2[i] *= invstd[L.lane * V + i];
    }
    // tree sum over the thread rows of each lane (sm: [blockDim][2V]; aliases the streaming slots)
    __syncthreads();
    float* mine = sm + threadIdx.x * (2 * V);
#pragma unroll
    for (int i = 0; i < V; ++i) { mine[i] = s1[i]; mine[V + i] = s2[i]; }
    int h = 1;
    while (h < L.R) h <<= 1;
    int rows = L.R;
    for (h >>= 1; h >= 1; h >>= 1) {
        __syncthreads();
        if (L.active && L.r < h && L.r + h < rows) {
            const float* o = sm + ((L.r + h) * L.lanes + L.lane) * (2 * V);
#pragma unroll
            for (int i = 0; i < V; ++i) { s1[i] += o[i]; s2[i] += o[V + i]; mine[i] = s1[i]; mine[V + i] = s2[i]; }
        }
        rows = rows < h ? rows : h;
    }
    if (L.active && L.r == 0) {
        float* out = partials + (long long)blockIdx.x * 2 * Cp;
#pragma unroll
        for (int i = 0; i < V; ++i) { out[L.lane * V + i] = s1[i]; out[Cp + L.lane * V + i] = s2[i]; }
    }
}

template <typename T>
__device__ __forceinline__ void bwd_reduce_to_partials(const T* __restrict__ g, const T* __restrict__ a,
                                                       const T* __restrict__ y, float* __restrict__ partials,
                                                       long long P, int Cp, int ld_g, int ld_a, int ld_y,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       int relu, float* sm) {
    if (relu) bwd_reduce_to_partials_t<T, 1>(g, a, y, partials, P, Cp, ld_g, ld_a, ld_y, mean, invstd, sm);
    else bwd_reduce_to_partials_t<T, 0>(g, a, y, partials, P, Cp, ld_g, ld_a, ld_y, mean, invstd, sm);
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y,
                     float* __restrict__ partials, long long P, int Cp, int ld_g, int ld_a, int ld_y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu) {
    extern __shared__ __align__(16) float sm[];  // [BN_THREADS][2*V]
    bwd_reduce_to_partials<T>(g, a, y, partials, P, Cp, ld_g, ld_a, ld_y, mean, invstd, relu, sm);
}

// sums[2][Cp] = sum over partials (one warp per channel)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int n_parts, int C, int Cp,
                                       float* __restrict__ sums) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= Cp) return;
    const int lane = threadIdx.x & 31;
    float a = 0.f, b = 0.f;
    if (c < C)
        for (int k = lane; k < n_parts; k += 32) { a += partials[(long long)k * 2 * Cp + c]; b += partials[(long long)k * 2 * Cp + Cp + c]; }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { sums[c] = a; sums[Cp + c] = b; }
}

// parameter grads from the LOCAL sums, normalisation coefficients from the (all-reduced) GLOBAL sums
__global__ void bn_bwd_coeffs_kernel(const float* __restrict__ sums, int C, int Cp, float inv_count,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                     const float* __restrict__ sums_local, float* __restrict__ c1, float* __restrict__ c2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    c1[c] = sums[c] * inv_count;
    c2[c] = sums[Cp + c] * inv_count;
    if (c < C) {
        const float db = sums_local[c], dg = sums_local[Cp + c];
        if (dbeta != nullptr) dbeta[c] = accumulate ? dbeta[c] + db : db;
        if (dgamma != nullptr) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    }
}

// backward pass 2: dy = scale * (dyb - c1 - xhat*c2);  dres (=|+=) dyb
template <typename T, int RELU, bool DRES>
__device__ __forceinline__ void bwd_elemt_pass_t(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y,
                                               T* __restrict__ dy, T* __restrict__ dres, long long P, int ld_g, int ld_a,
                                               int ld_y, int ld_dy, int ld_dres, const float* __restrict__ mean,
                                               const float* __restrict__ invstd, const float* __restrict__ scale,
                                               const float* __restrict__ c1, const float* __restrict__ c2,
                                               int acc_dy, int acc_dres, const Lay<T>& L, float* sm, bool reverse,
                                               const float* __restrict__ shift = nullptr) {
    constexpr int V = Vec<T>::N;
    constexpr int relu = RELU;
    if (!DRES) dres = nullptr;
    if (!L.active) return;
    const int c0 = L.lane * V;
    float mu[V], k2[V], sc[V], k1[V], sh[V];      // dy = sc * (g - k1 - (y - mu) * k2),  k2 = invstd * c2
#pragma unroll
    for (int k = 0; k < V; ++k) {
        mu[k] = mean[c0 + k]; k2[k] = invstd[c0 + k] * c2[c0 + k]; sc[k] = scale[c0 + k]; k1[k] = c1[c0 + k];
    }
    if (relu == 2) {
#pragma unroll
        for (int k = 0; k < V; ++k) sh[k] = shift[c0 + k];
    }
    auto one = [&](long long q, const uint4& gv, const uint4& yv, const uint4& av) {
        uint4 od, orr;
        if (acc_dy) od = ld16<T>(dy + q * ld_dy + c0);
        if (dres != nullptr && acc_dres) orr = ld16<T>(dres + q * ld_dres + c0);
        Vec<T> o, gm;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            float gg = elem<T>(gv, k);
            const float yy = elem<T>(yv, k);
            if (relu == 1) gg = elem<T>(av, k) > 0.f ? gg : 0.f;
            if (relu == 2) gg = fmaf(yy, sc[k], sh[k]) > 0.f ? gg : 0.f;
            gm.v[k] = gg;
            o.v[k] = sc[k] * (gg - k1[k] - (yy - mu[k]) * k2[k]);
        }
        if (acc_dy) {
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] += elem<T>(od, k);
        }
        o.store(dy + q * ld_dy + c0);
        if (dres != nullptr) {
            if (acc_dres) {
#pragma unroll
                for (int k = 0; k < V; ++k) gm.v[k] += elem<T>(orr, k);
            }
            gm.store(dres + q * ld_dres + c0);
        }
    };
    if (relu == 1) {
        const T* const base[3] = {g, y, a};
        const int lds[3] = {ld_g, ld_y, ld_a};
        stream_pixels<T, 3>(sm, L, P, reverse, base, lds, [&](long long q, const uint4 (&v)[3]) { one(q, v[0], v[1], v[2]); });
    } else {
        const T* const base[2] = {g, y};
        const int lds[2] = {ld_g, ld_y};
        stream_pixels<T, 2>(sm, L, P, reverse, base, lds, [&](long long q, const uint4 (&v)[2]) { one(q, v[0], v[1], v[1]); });
    }
}

template <typename T>
__device__ __forceinline__ void bwd_elemt_pass(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y,
                                               T* __restrict__ dy, T* __restrict__ dres, long long P, int ld_g, int ld_a,
                                               int ld_y, int ld_dy, int ld_dres, const float* __restrict__ mean,
                                               const float* __restrict__ invstd, const float* __restrict__ scale,
                                               const float* __restrict__ c1, const float* __restrict__ c2, int relu,
                                               int acc_dy, int acc_dres, const Lay<T>& L, float* sm) {
    if (relu) bwd_elemt_pass_t<T, 1, true>(g, a, y, dy, dres, P, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1,
                                           c2, acc_dy, acc_dres, L, sm, false);
    else bwd_elemt_pass_t<T, 0, true>(g, a, y, dy, dres, P, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2,
                                      acc_dy, acc_dres, L, sm, false);
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_elemt_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y, T* __restrict__ dy,
                    T* __restrict__ dres, long long P, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
                    const float* __restrict__ c1, const float* __restrict__ c2, int relu, int acc_dy, int acc_dres) {
    extern __shared__ __align__(16) float sm[];
    const Lay<T> L(Cp);
    bwd_elemt_pass<T>(g, a, y, dy, dres, P, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2, relu, acc_dy,
                      acc_dres, L, sm);
}

// ---------------------------------------------------------------------------
// fused single-rank training BN: one cooperative launch per direction
// ---------------------------------------------------------------------------
// optional phase timestamps (globaltimer ns) of CTA 0, for tools/microbench_bn.py; null in production
static unsigned long long* g_bn_prof = nullptr;
void bn_debug_set_prof(unsigned long long* p) { g_bn_prof = p; }
__device__ __forceinline__ void prof_mark(unsigned long long* prof, int i) {
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prof[i] = t;
    }
}

// ---- SyncBN over NVLink peer memory (phase 3) --------------------------------------------------------------------------
// One cooperative launch per BN, as on a single GPU: after the rank-local merge the CTA that finalizes channel c stores its
// (count, mean, M2) -- or (sum dy, sum dy*xhat) in the backward -- straight into every rank's mailbox (P2P stores through
// NVSwitch), then reads the other ranks' values from its OWN mailbox and merges them in rank order, so every rank computes
// bit-identical statistics.  Each value travels as one 8-byte word {float bits, sequence number}: an aligned 8-byte store
// is single-copy atomic, so the flag can never be seen ahead of its data and no fence is needed (NCCL's LL protocol).  The
// sequence number comes from a per-op device counter (the launch arguments are frozen inside CUDA graphs); words of
// consecutive launches of an op alternate between two buffers.  A wait that exceeds kPeerSpinLimit (~60 s) sets a sticky error flag
// and gives up (the host raises at the end of the step) instead of hanging the GPU.
constexpr int kMaxPeers = 8;
constexpr long long kPeerSpinLimit = 120000000000LL;    // ~60 s of SM clocks (ranks record their plans at different speeds)

struct BnPeer {
    int world, rank;
    unsigned long long* box[kMaxPeers];   // this op's mailbox slot on rank r (peer-mapped); box[rank] is local
    unsigned int* seq;                    // launches of this op so far (local)
    int* err;                             // sticky (local)
};

__device__ __forceinline__ void peer_put(unsigned long long* p, float v, unsigned int seq) {
    const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}

__device__ __forceinline__ float peer_get(const unsigned long long* p, unsigned int seq, int* err) {
    unsigned long long w;
    const long long t0 = clock64();
    int spins = 0;
    while (true) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        if ((unsigned int)(w >> 32) == seq) break;
        if ((++spins & 255) == 0) {
            if (*(volatile int*)err != 0) break;
            if (clock64() - t0 > kPeerSpinLimit) { *(volatile int*)err = 1; break; }
        }
    }
    return __uint_as_float((unsigned int)w);
}

// word index of value k (of nv) of channel c, group gq, written by rank src, buffer par
__device__ __forceinline__ long long peer_word(int par, int world, int src, int G, int gq, int nv, int k, int Cp, int c) {
    return ((((long long)par * world + src) * G + gq) * nv + k) * Cp + c;
}

struct BnFwdArgs {
    unsigned long long* prof;
    const void *y, *res;
    void* out;
    float* partials;
    long long P;
    int C, Cp, ld_y, ld_res, ld_out, relu;
    const float *gamma, *beta;
    float *running_mean, *running_var;
    long long* nbt;
    float momentum, eps;
    float *mean, *invstd, *scale, *shift;
    // statistics groups (stacked calls of the reference along the batch axis): group g covers P pixels starting
    // gs_* elements after group g-1, its statistics outputs sit stat_stride floats after the previous group's.  The
    // grid is G x (gridDim.x / G) CTAs; running statistics take the G momentum updates in group order.
    int G;
    long long gs_y, gs_res, gs_out;
    int stat_stride;
    // SyncBN halves (a collective sits between them): phase 0 = the whole BN; 1 = statistics + rank-local merge only,
    // group g's (count, mean, M2) goes to msg + g*3*Cp; 2 = finalize from `ext_parts` gathered partial sets (set r of
    // group g at ext + r*ext_stride + g*3*Cp) + apply.
    int phase;
    float* msg;
    const float* ext;
    int ext_parts;
    long long ext_stride;
    BnPeer peer;                                   // phase 3: the whole BN with the cross-rank merge inside (see BnPeer)
};

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_fwd_fused_kernel(const BnFwdArgs A) {
    extern __shared__ __align__(16) float sm[];
    cg::grid_group grid = cg::this_grid();
    prof_mark(A.prof, 0);
    const int nbpg = gridDim.x / A.G;                 // CTAs per statistics group
    const int grp = blockIdx.x / nbpg, lb = blockIdx.x - grp * nbpg;
    const unsigned int pseq = A.phase == 3 ? *(volatile unsigned int*)A.peer.seq + 1u : 0u;   // bumped after the last grid.sync
    // parameters of the channel this CTA will finalize: cold in HBM, so fetched under the statistics pass
    float gm0 = 0.f, bt0 = 0.f, rm00 = 0.f, rv00 = 0.f;
    if (A.phase != 1 && threadIdx.x == 0 && (int)blockIdx.x < A.C) {
        gm0 = A.gamma[blockIdx.x]; bt0 = A.beta[blockIdx.x];
        if (A.running_mean != nullptr) { rm00 = A.running_mean[blockIdx.x]; rv00 = A.running_var[blockIdx.x]; }
    }
    if (A.phase != 2) {
        stats_to_partials<T>((const T*)A.y + grp * A.gs_y, A.partials, A.P, A.Cp, A.ld_y, sm, lb, nbpg);
        prof_mark(A.prof, 1);
        grid.sync();
    }
    prof_mark(A.prof, 2);
    // One CTA per channel: every thread fetches <= 3 partials at once (one L2 round trip for the whole merge),
    // warp butterflies, then the 8 warp results meet in shared memory.  Groups are finalized one after the other by
    // the same CTA so that the running statistics see their momentum updates in the reference's call order.
    for (int c = blockIdx.x; c < A.Cp; c += gridDim.x) {
        const int n_parts = A.phase == 2 ? A.ext_parts : nbpg, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const long long part_stride = A.phase == 2 ? A.ext_stride : 3LL * A.Cp;
        float gm = gm0, bt = bt0, rm0 = rm00, rv0 = rv00;      // first channel: fetched before the statistics pass
        if (A.phase != 1 && threadIdx.x == 0 && c < A.C && c != blockIdx.x) {
            gm = A.gamma[c]; bt = A.beta[c];
            if (A.running_mean != nullptr) { rm0 = A.running_mean[c]; rv0 = A.running_var[c]; }
        }
        for (int gq = 0; gq < A.G; ++gq) {
            const float* parts = A.phase == 2 ? A.ext + (long long)gq * 3 * A.Cp : A.partials + (long long)gq * nbpg * 3 * A.Cp;
            float n = 0.f, mean = 0.f, m2 = 0.f;
            if (c < A.C) {
                float a[3][3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int k = threadIdx.x + j * BN_THREADS;
                    const bool valid = k < n_parts;
                    const float* p = parts + (long long)(valid ? k : 0) * part_stride;
                    a[j][0] = valid ? __ldcg(p + c) : 0.f;
                    a[j][1] = __ldcg(p + A.Cp + c);
                    a[j][2] = __ldcg(p + 2 * A.Cp + c);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) chan_merge(n, mean, m2, a[j][0], a[j][1], a[j][2]);
                warp_tree_merge(n, mean, m2);
            }
            __syncthreads();
            if (lane == 0) { sm[warp * 3] = n; sm[warp * 3 + 1] = mean; sm[warp * 3 + 2] = m2; }
            __syncthreads();
            if (warp == 0) {
                const int w = lane & 7;
                n = sm[w * 3]; mean = sm[w * 3 + 1]; m2 = sm[w * 3 + 2];
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {      // butterfly over the 8 warp results (lanes 8.. mirror lanes 0..7)
                    const float nb = __shfl_xor_sync(0xffffffffu, n, o);
                    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
                    const float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
                    const float nn = n + nb;
                    if (nn > 0.f) {
                        const float d = mb - mean;
                        const float f = nb / nn;
                        mean = (nb > 0.f && n > 0.f) ? mean + d * f : (nb > 0.f ? mb : mean);
                        m2 = m2 + m2b + ((nb > 0.f && n > 0.f) ? d * d * n * f : 0.f);
                    }
                    n = nn;
                }
                if (A.phase == 3 && c < A.C) {                 // cross-rank merge through peer memory (warp 0, lane r <-> rank r)
                    const BnPeer& X = A.peer;
                    const int par = (int)(pseq & 1u);
                    n = __shfl_sync(0xffffffffu, n, 0); mean = __shfl_sync(0xffffffffu, mean, 0); m2 = __shfl_sync(0xffffffffu, m2, 0);
                    float rn = 0.f, rmean = 0.f, rm2 = 0.f;
                    if (lane < X.world) {
                        unsigned long long* dst = X.box[lane] + peer_word(par, X.world, X.rank, A.G, gq, 3, 0, A.Cp, c);
                        peer_put(dst, n, pseq); peer_put(dst + A.Cp, mean, pseq); peer_put(dst + 2 * A.Cp, m2, pseq);
                        const unsigned long long* src = X.box[X.rank] + peer_word(par, X.world, lane, A.G, gq, 3, 0, A.Cp, c);
                        rn = peer_get(src, pseq, X.err); rmean = peer_get(src + A.Cp, pseq, X.err); rm2 = peer_get(src + 2 * A.Cp, pseq, X.err);
                    }
                    n = 0.f; mean = 0.f; m2 = 0.f;
                    for (int r = 0; r < X.world; ++r)          // fixed rank order: identical on every rank
                        chan_merge(n, mean, m2, __shfl_sync(0xffffffffu, rn, r), __shfl_sync(0xffffffffu, rmean, r),
                                   __shfl_sync(0xffffffffu, rm2, r));
                }
                if (lane == 0 && A.phase == 1) {               // rank-local (count, mean, M2) of this group -> the message
                    float* o = A.msg + (long long)gq * 3 * A.Cp;
                    o[c] = c < A.C ? n : 0.f; o[A.Cp + c] = c < A.C ? mean : 0.f; o[2 * A.Cp + c] = c < A.C ? m2 : 0.f;
                } else if (lane == 0) {
                    const int so = gq * A.stat_stride + c;
                    if (c >= A.C) {
                        A.mean[so] = 0.f; A.invstd[so] = 0.f; A.scale[so] = 0.f; A.shift[so] = 0.f;
                    } else {
                        const float var = m2 / n;
                        const float invstd = rsqrtf(var + A.eps);
                        const float sc = gm * invstd;
                        A.mean[so] = mean; A.invstd[so] = invstd; A.scale[so] = sc; A.shift[so] = bt - mean * sc;
                        if (A.running_mean != nullptr) {
                            const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
                            rm0 = (1.f - A.momentum) * rm0 + A.momentum * mean;
                            rv0 = (1.f - A.momentum) * rv0 + A.momentum * unbiased;
                            if (gq == A.G - 1) { A.running_mean[c] = rm0; A.running_var[c] = rv0; }
                        }
                    }
                }
            }
        }
    }
    if (A.phase == 1) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.nbt != nullptr) *A.nbt += A.G;
    prof_mark(A.prof, 3);
    grid.sync();
    prof_mark(A.prof, 4);
    if (A.phase == 3 && blockIdx.x == 0 && threadIdx.x == 0) *A.peer.seq = pseq;
    const Lay<T> L(A.Cp, lb, nbpg);
    apply_pass<T>((const T*)A.y + grp * A.gs_y, A.res != nullptr ? (const T*)A.res + grp * A.gs_res : nullptr,
                  (T*)A.out + grp * A.gs_out, A.P, A.ld_y, A.ld_res, A.ld_out, A.scale + grp * A.stat_stride,
                  A.shift + grp * A.stat_stride, A.relu, L, sm, true);
    prof_mark(A.prof, 5);
}

struct BnBwdArgs {
    unsigned long long* prof;
    const void *g, *a, *y;
    void *dy, *dres;
    float* partials;
    long long P;
    int C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, relu, acc_dy, acc_dres, acc_param;
    const float *mean, *invstd, *scale, *shift;
    float *dgamma, *dbeta, *c1, *c2;
    float inv_count;
    int G;                                         // statistics groups, as in BnFwdArgs; d(gamma), d(beta) sum over them
    long long gs_g, gs_a, gs_y, gs_dy, gs_dres;
    int stat_stride;
    // SyncBN halves: phase 1 = reduce + per-channel rank-local sums (group g's (sum dyb, sum dyb*xhat) to msg + g*2*Cp,
    // d(gamma)/d(beta) from those LOCAL sums, as torch's SyncBatchNorm does); phase 2 = coefficients from the all-reduced
    // sums at ext + g*2*Cp (inv_count covers every rank's pixels) + the elementwise pass.
    int phase;
    float* msg;
    const float* ext;
    BnPeer peer;                                   // phase 3: sums exchanged through peer memory inside the launch
};

template <typename T, int RELU, bool DRES>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_fused_kernel(const BnBwdArgs A) {
    extern __shared__ __align__(16) float sm[];
    cg::grid_group grid = cg::this_grid();
    prof_mark(A.prof, 0);
    const int nbpg = gridDim.x / A.G;
    const int grp = blockIdx.x / nbpg, lb = blockIdx.x - grp * nbpg;
    const int so_g = grp * A.stat_stride;
    const T* gp = (const T*)A.g + grp * A.gs_g;
    const T* ap = A.a != nullptr ? (const T*)A.a + grp * A.gs_a : nullptr;
    const T* yp = (const T*)A.y + grp * A.gs_y;
    const unsigned int pseq = A.phase == 3 ? *(volatile unsigned int*)A.peer.seq + 1u : 0u;
    if (A.phase != 2) {
        bwd_reduce_to_partials_t<T, RELU>(gp, ap, yp, A.partials, A.P, A.Cp, A.ld_g, A.ld_a, A.ld_y, A.mean + so_g,
                                          A.invstd + so_g, sm, A.scale + so_g, A.shift != nullptr ? A.shift + so_g : nullptr,
                                          lb, nbpg);
        prof_mark(A.prof, 1);
        grid.sync();
    }
    prof_mark(A.prof, 2);
    for (int c = blockIdx.x; c < A.Cp; c += gridDim.x) {      // one CTA per channel, one L2 round trip per group
        const int n_parts = nbpg, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float t1 = 0.f, t2 = 0.f;                              // sums over the groups: parameter gradients
        for (int gq = 0; gq < A.G; ++gq) {
            if (A.phase == 2) {                                // global sums arrive all-reduced: one value per channel
                if (threadIdx.x == 0) {
                    const float* e = A.ext + (long long)gq * 2 * A.Cp;
                    A.c1[gq * A.stat_stride + c] = (c < A.C ? e[c] : 0.f) * A.inv_count;
                    A.c2[gq * A.stat_stride + c] = (c < A.C ? e[A.Cp + c] : 0.f) * A.inv_count;
                }
                continue;
            }
            const float* parts = A.partials + (long long)gq * nbpg * 2 * A.Cp;
            float s1 = 0.f, s2 = 0.f;
            if (c < A.C) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int k = threadIdx.x + j * BN_THREADS;
                    if (k < n_parts) {
                        s1 += __ldcg(parts + (long long)k * 2 * A.Cp + c);
                        s2 += __ldcg(parts + (long long)k * 2 * A.Cp + A.Cp + c);
                    }
                }
            }
            s1 = warp_sum(s1); s2 = warp_sum(s2);
            __syncthreads();
            if (lane == 0) { sm[warp * 2] = s1; sm[warp * 2 + 1] = s2; }
            __syncthreads();
            if (warp == 0) {
                s1 = 0.f; s2 = 0.f;
#pragma unroll
                for (int w = 0; w < BN_THREADS / 32; ++w) { s1 += sm[w * 2]; s2 += sm[w * 2 + 1]; }     // every lane: the local sums
                float g1 = s1, g2 = s2;
                if (A.phase == 3 && c < A.C) {                 // all-reduce through peer memory (lane r <-> rank r), rank order
                    const BnPeer& X = A.peer;
                    const int par = (int)(pseq & 1u);
                    float r1 = 0.f, r2 = 0.f;
                    if (lane < X.world) {
                        unsigned long long* dst = X.box[lane] + peer_word(par, X.world, X.rank, A.G, gq, 2, 0, A.Cp, c);
                        peer_put(dst, s1, pseq); peer_put(dst + A.Cp, s2, pseq);
                        const unsigned long long* src = X.box[X.rank] + peer_word(par, X.world, lane, A.G, gq, 2, 0, A.Cp, c);
                        r1 = peer_get(src, pseq, X.err); r2 = peer_get(src + A.Cp, pseq, X.err);
                    }
                    g1 = 0.f; g2 = 0.f;
                    for (int r = 0; r < X.world; ++r) { g1 += __shfl_sync(0xffffffffu, r1, r); g2 += __shfl_sync(0xffffffffu, r2, r); }
                }
                if (lane == 0) {
                    if (A.phase == 1) {
                        float* o = A.msg + (long long)gq * 2 * A.Cp;
                        o[c] = s1; o[A.Cp + c] = s2;
                    } else {
                        A.c1[gq * A.stat_stride + c] = g1 * A.inv_count;
                        A.c2[gq * A.stat_stride + c] = g2 * A.inv_count;
                    }
                    t1 += s1; t2 += s2;                        // parameter gradients from the LOCAL sums (DDP reduces them)
                }
            }
        }
        if (A.phase == 2) continue;
        if (threadIdx.x == 0 && c < A.C) {
            if (A.dbeta != nullptr) A.dbeta[c] = A.acc_param ? A.dbeta[c] + t1 : t1;
            if (A.dgamma != nullptr) A.dgamma[c] = A.acc_param ? A.dgamma[c] + t2 : t2;
        }
    }
    if (A.phase == 1) return;
    prof_mark(A.prof, 3);
    grid.sync();
    prof_mark(A.prof, 4);
    if (A.phase == 3 && blockIdx.x == 0 && threadIdx.x == 0) *A.peer.seq = pseq;
    const Lay<T> L(A.Cp, lb, nbpg);
    bwd_elemt_pass_t<T, RELU, DRES>(gp, ap, yp, (T*)A.dy + grp * A.gs_dy, A.dres != nullptr ? (T*)A.dres + grp * A.gs_dres : nullptr,
                                    A.P, A.ld_g, A.ld_a, A.ld_y, A.ld_dy, A.ld_dres, A.mean + so_g, A.invstd + so_g,
                                    A.scale + so_g, A.c1 + so_g, A.c2 + so_g, A.acc_dy, A.acc_dres, L, sm, true,
                                    A.shift != nullptr ? A.shift + so_g : nullptr);
    prof_mark(A.prof, 5);
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline int vec_of(int dtype) { return dtype == VAE2_DT_F32 ? 4 : 8; }

static int reduce_grid(long long P, int Cp, int V) {
    const int lanes = Cp / V;
    const int R = BN_THREADS / lanes;
    long long need = (P + (long long)R * 2 - 1) / ((long long)R * 2);   // 2 pixels per thread until the CTA cap, then more
    if (need < 1) need = 1;
    if (need > BN_SPLIT_PARTS) need = BN_SPLIT_PARTS;
    return (int)need;
}

// elementwise passes: >= 4 pixels per thread, at most 8 CTAs per SM
static int elemt_grid(long long P, int Cp, int V) {
    const int R = BN_THREADS / (Cp / V);
    long long need = (P + (long long)R * 4 - 1) / ((long long)R * 4);
    if (need < 1) need = 1;
    if (need > kNumSMs * 8) need = kNumSMs * 8;
    return (int)need;
}

// Cooperative grid: every CTA must be resident, so the size is capped by the occupancy the driver reports.
template <typename K>
static int coop_grid(K kernel, size_t smem, long long P, int Cp, int V, int* cache) {
    if (*cache == 0) {
        int per_sm = 0, sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BN_THREADS, smem) != cudaSuccess || per_sm < 1)
            return 0;
        *cache = per_sm * sms;
    }
    const int R = BN_THREADS / (Cp / V);
    long long need = (P + (long long)R * 4 - 1) / ((long long)R * 4);
    if (need < 1) need = 1;
    if (need > BN_MAX_PARTS) need = BN_MAX_PARTS;
    if (need > *cache) need = *cache;
    static int env_cap = -1;
    if (env_cap < 0) { const char* e = getenv("VAE2_BN_MAXGRID"); env_cap = e ? atoi(e) : 0; }
    if (env_cap > 0 && need > env_cap) need = env_cap;
    return (int)need;
}

template <typename T>
static int launch_fwd_fused(BnFwdArgs& A, cudaStream_t st) {
    constexpr int V = Vec<T>::N;
    static int cap = 0;
    const size_t smem = kPipeBytes;   // streaming slots; the merge scratch aliases them
    int grid = coop_grid(bn_fwd_fused_kernel<T>, smem, A.P * A.G, A.Cp, V, &cap);
    if (grid < 1) return VAE2_ERR_CUDA;
    grid = grid < A.G ? A.G : grid / A.G * A.G;      // the same CTA count for every statistics group (cap >= 148 > G)
    void* args[] = {(void*)&A};
    if (cudaLaunchCooperativeKernel((const void*)bn_fwd_fused_kernel<T>, dim3(grid), dim3(BN_THREADS), args, smem, st) !=
        cudaSuccess)
        return VAE2_ERR_CUDA;
    return check_launch();
}

template <typename T, int RELU, bool DRES>
static int launch_bwd_fused_t(BnBwdArgs& A, cudaStream_t st) {
    constexpr int V = Vec<T>::N;
    static int cap = 0;
    const size_t smem = kPipeBytes;   // streaming slots; the merge scratch aliases them
    int grid = coop_grid(bn_bwd_fused_kernel<T, RELU, DRES>, smem, A.P * A.G, A.Cp, V, &cap);
    if (grid < 1) return VAE2_ERR_CUDA;
    grid = grid < A.G ? A.G : grid / A.G * A.G;
    void* args[] = {(void*)&A};
    if (cudaLaunchCooperativeKernel((const void*)bn_bwd_fused_kernel<T, RELU, DRES>, dim3(grid), dim3(BN_THREADS), args,
                                    smem, st) != cudaSuccess)
        return VAE2_ERR_CUDA;
    return check_launch();
}

template <typename T>
static int launch_bwd_fused(BnBwdArgs& A, cudaStream_t st) {
    const bool dres = A.dres != nullptr;
    switch (A.relu * 2 + (dres ? 1 : 0)) {
        case 0: return launch_bwd_fused_t<T, 0, false>(A, st);
        case 1: return launch_bwd_fused_t<T, 0, true>(A, st);
        case 2: return launch_bwd_fused_t<T, 1, false>(A, st);
        case 3: return launch_bwd_fused_t<T, 1, true>(A, st);
        case 4: return launch_bwd_fused_t<T, 2, false>(A, st);
        default: return launch_bwd_fused_t<T, 2, true>(A, st);
    }
}

int bn_fwd_fused(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                 int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale,
                 float* shift, int relu, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || ld_out % V || (res && ld_res % V) || Cp / V > BN_THREADS || P < 1) return VAE2_ERR_ARG;
    return bn_fwd_fused_groups(y, res, out, partials, dtype, P, C, Cp, ld_y, ld_res, ld_out, gamma, beta, running_mean,
                               running_var, nbt, momentum, eps, mean, invstd, scale, shift, relu, 1, 0, st);
}

int bn_fwd_fused_groups(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                        int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd,
                        float* scale, float* shift, int relu, int groups, int stat_stride, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || ld_out % V || (res && ld_res % V) || Cp / V > BN_THREADS || P < 1 || groups < 1 ||
        groups > 64)
        return VAE2_ERR_ARG;
    BnFwdArgs A{g_bn_prof, y, res, out, partials, P, C, Cp, ld_y, ld_res, ld_out, relu, gamma, beta, running_mean, running_var,
                nbt, momentum, eps, mean, invstd, scale, shift, groups, P * ld_y, P * ld_res, P * ld_out, stat_stride,
                0, nullptr, nullptr, 0, 0};
    return dtype == VAE2_DT_F32 ? launch_fwd_fused<float>(A, st) : launch_fwd_fused<__nv_bfloat16>(A, st);
}

// SyncBN forward, first half: statistics of every group + rank-local merge -> msg[g][3][Cp] (one cooperative launch)
int bn_sync_fwd_stats(const void* y, float* partials, int dtype, long long P, int C, int Cp, int ld_y, int groups, float* msg,
                      cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || Cp / V > BN_THREADS || P < 1 || groups < 1 || groups > 64) return VAE2_ERR_ARG;
    BnFwdArgs A{g_bn_prof, y, nullptr, nullptr, partials, P, C, Cp, ld_y, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                0.f, 0.f, nullptr, nullptr, nullptr, nullptr, groups, P * ld_y, 0, 0, 0, 1, msg, nullptr, 0, 0};
    return dtype == VAE2_DT_F32 ? launch_fwd_fused<float>(A, st) : launch_fwd_fused<__nv_bfloat16>(A, st);
}

// SyncBN forward, second half: finalize every group from `n_parts` gathered sets + running statistics + apply
int bn_sync_fwd_apply(const void* y, const void* res, void* out, int dtype, long long P, int C, int Cp, int ld_y, int ld_res,
                      int ld_out, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                      int relu, int groups, int stat_stride, const float* gathered, int n_parts, long long part_stride,
                      cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || ld_out % V || (res && ld_res % V) || Cp / V > BN_THREADS || P < 1 || groups < 1 ||
        groups > 64 || n_parts < 1 || n_parts > 3 * BN_THREADS)
        return VAE2_ERR_ARG;
    BnFwdArgs A{g_bn_prof, y, res, out, nullptr, P, C, Cp, ld_y, ld_res, ld_out, relu, gamma, beta, running_mean, running_var,
                nbt, momentum, eps, mean, invstd, scale, shift, groups, P * ld_y, P * ld_res, P * ld_out, stat_stride,
                2, nullptr, gathered, n_parts, part_stride};
    return dtype == VAE2_DT_F32 ? launch_fwd_fused<float>(A, st) : launch_fwd_fused<__nv_bfloat16>(A, st);
}

int bn_bwd_fused(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                 long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                 const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                 int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || ld_dy % V || (relu == 1 && (a == nullptr || ld_a % V)) ||
        (relu == 2 && shift == nullptr) || relu < 0 || relu > 2 || (dres && ld_dres % V) || Cp / V > BN_THREADS || P < 1)
        return VAE2_ERR_ARG;
    return bn_bwd_fused_groups(g, a, y, dy, dres, partials, dtype, P, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd,
                               scale, shift, dgamma, dbeta, accumulate_param, c1, c2, relu, acc_dy, acc_dres, 1, 0, st);
}

int bn_bwd_fused_groups(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                        long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                        const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                        int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                        int stat_stride, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || ld_dy % V || (relu == 1 && (a == nullptr || ld_a % V)) ||
        (relu == 2 && shift == nullptr) || relu < 0 || relu > 2 || (dres && ld_dres % V) || Cp / V > BN_THREADS || P < 1 ||
        groups < 1 || groups > 64)
        return VAE2_ERR_ARG;
    BnBwdArgs A{g_bn_prof, g, a, y, dy, dres, partials, P, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, relu, acc_dy, acc_dres,
                accumulate_param, mean, invstd, scale, shift, dgamma, dbeta, c1, c2, 1.0f / (float)P, groups, P * ld_g,
                P * ld_a, P * ld_y, P * ld_dy, P * ld_dres, stat_stride, 0, nullptr, nullptr};
    return dtype == VAE2_DT_F32 ? launch_bwd_fused<float>(A, st) : launch_bwd_fused<__nv_bfloat16>(A, st);
}

// SyncBN backward halves (phase 1: reduce + local sums -> msg[g][2][Cp], d(gamma)/d(beta); phase 2: coefficients from the
// all-reduced sums `gsum` with inv_count over ALL ranks' pixels + elementwise pass).  `dres` only matters in phase 2, but
// both halves must name the same kernel variant, so it is passed to both.
int bn_sync_bwd(int phase, const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups, int stat_stride,
                float* msg, const float* gsum, float inv_count, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || (phase == 2 && ld_dy % V) || (relu == 1 && (a == nullptr || ld_a % V)) ||
        (relu == 2 && shift == nullptr) || relu < 0 || relu > 2 || (dres && ld_dres % V) || Cp / V > BN_THREADS || P < 1 ||
        groups < 1 || groups > 64 || (phase != 1 && phase != 2))
        return VAE2_ERR_ARG;
    BnBwdArgs A{g_bn_prof, g, a, y, dy, dres, partials, P, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, relu, acc_dy, acc_dres,
                accumulate_param, mean, invstd, scale, shift, dgamma, dbeta, c1, c2, inv_count, groups, P * ld_g,
                P * ld_a, P * ld_y, P * ld_dy, P * ld_dres, stat_stride, phase, msg, gsum};
    return dtype == VAE2_DT_F32 ? launch_bwd_fused<float>(A, st) : launch_bwd_fused<__nv_bfloat16>(A, st);
}

// ---- SyncBN over peer memory: process-wide context + the two launches ---------------------------------------------------
static struct {
    int world, rank;
    unsigned long long* base[kMaxPeers];
    unsigned int* seq;
    int* err;
} g_peer = {0, 0, {nullptr}, nullptr, nullptr};

// bases[r]: rank r's mailbox as mapped into THIS process (bases[rank] is the local allocation); seq: local zeroed counters,
// one per op; err: local zeroed int.  world = 0 tears the context down.
int bn_peer_setup(int world, int rank, void* const* bases, unsigned int* seq, int* err) {
    if (world == 0) { g_peer.world = 0; return VAE2_OK; }
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || bases == nullptr || seq == nullptr || err == nullptr)
        return VAE2_ERR_ARG;
    g_peer.world = world; g_peer.rank = rank; g_peer.seq = seq; g_peer.err = err;
    for (int r = 0; r < world; ++r) g_peer.base[r] = reinterpret_cast<unsigned long long*>(bases[r]);
    return VAE2_OK;
}

// 8-byte words one op needs in every rank's mailbox: 2 buffers x world senders x groups x values x lanes
long long bn_peer_slot_words(int world, int groups, int Cp, int backward) {
    return 2LL * world * groups * (backward ? 2 : 3) * Cp;
}

static int fill_peer(BnPeer& X, long long slot_word, int seq_index) {
    if (g_peer.world < 1 || slot_word < 0 || seq_index < 0) return VAE2_ERR_ARG;
    X.world = g_peer.world; X.rank = g_peer.rank;
    for (int r = 0; r < kMaxPeers; ++r) X.box[r] = r < g_peer.world ? g_peer.base[r] + slot_word : nullptr;
    X.seq = g_peer.seq + seq_index; X.err = g_peer.err;
    return VAE2_OK;
}

// SyncBN forward in ONE cooperative launch: statistics, cross-rank merge through peer memory, running statistics, apply.
// P = this rank's pixels per group; the merged count covers every rank.
int bn_fwd_fused_peer(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                      int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd,
                      float* scale, float* shift, int relu, int groups, int stat_stride, long long slot_word, int seq_index,
                      cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || ld_out % V || (res && ld_res % V) || Cp / V > BN_THREADS || P < 1 || groups < 1 ||
        groups > 64)
        return VAE2_ERR_ARG;
    BnFwdArgs A{g_bn_prof, y, res, out, partials, P, C, Cp, ld_y, ld_res, ld_out, relu, gamma, beta, running_mean, running_var,
                nbt, momentum, eps, mean, invstd, scale, shift, groups, P * ld_y, P * ld_res, P * ld_out, stat_stride,
                3, nullptr, nullptr, 0, 0};
    if (int e = fill_peer(A.peer, slot_word, seq_index)) return e;
    return dtype == VAE2_DT_F32 ? launch_fwd_fused<float>(A, st) : launch_fwd_fused<__nv_bfloat16>(A, st);
}

// SyncBN backward in ONE cooperative launch; inv_count = 1 / (pixels of a group over ALL ranks); d(gamma), d(beta) come
// from the rank-local sums (DDP all-reduces parameter gradients), as in torch's SyncBatchNorm.
int bn_bwd_fused_peer(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                      long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                      const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                      int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                      int stat_stride, float inv_count, long long slot_word, int seq_index, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || ld_dy % V || (relu == 1 && (a == nullptr || ld_a % V)) ||
        (relu == 2 && shift == nullptr) || relu < 0 || relu > 2 || (dres && ld_dres % V) || Cp / V > BN_THREADS || P < 1 ||
        groups < 1 || groups > 64)
        return VAE2_ERR_ARG;
    BnBwdArgs A{g_bn_prof, g, a, y, dy, dres, partials, P, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, relu, acc_dy, acc_dres,
                accumulate_param, mean, invstd, scale, shift, dgamma, dbeta, c1, c2, inv_count, groups, P * ld_g,
                P * ld_a, P * ld_y, P * ld_dy, P * ld_dres, stat_stride, 3, nullptr, nullptr};
    if (int e = fill_peer(A.peer, slot_word, seq_index)) return e;
    return dtype == VAE2_DT_F32 ? launch_bwd_fused<float>(A, st) : launch_bwd_fused<__nv_bfloat16>(A, st);
}

// CUDA IPC plumbing for the mailboxes (one cudaMalloc per process, zeroed; peers map it with cudaIpcOpenMemHandle)
int ipc_alloc(long long bytes, void** ptr, void* handle64) {
    if (bytes < 1 || ptr == nullptr || handle64 == nullptr) return VAE2_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (cudaMalloc(ptr, (size_t)bytes) != cudaSuccess) return VAE2_ERR_CUDA;
    if (cudaMemset(*ptr, 0, (size_t)bytes) != cudaSuccess) return VAE2_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return VAE2_ERR_CUDA;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, *ptr) != cudaSuccess) { cudaGetLastError(); return VAE2_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    return VAE2_OK;
}
int ipc_open(const void* handle64, void** ptr) {
    if (ptr == nullptr || handle64 == nullptr) return VAE2_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return VAE2_ERR_CUDA; }
    return VAE2_OK;
}
int ipc_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? VAE2_OK : VAE2_ERR_CUDA; }
int ipc_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? VAE2_OK : VAE2_ERR_CUDA; }

int bn_stats(const void* y, float* partials, int* n_partials_out, int dtype, long long P, int Cp, int ld, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld % V || Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = reduce_grid(P, Cp, V);
    if (n_partials_out) *n_partials_out = grid;
    const size_t smem = kPipeBytes;   // streaming slots; the merge scratch aliases them
    if (dtype == VAE2_DT_F32)
        bn_stats_kernel<float><<<grid, BN_THREADS, smem, st>>>((const float*)y, partials, P, Cp, ld);
    else
        bn_stats_kernel<__nv_bfloat16><<<grid, BN_THREADS, smem, st>>>((const __nv_bfloat16*)y, partials, P, Cp, ld);
    return check_launch();
}

int bn_merge(const float* partials, int n_partials, int Cp, float* merged, cudaStream_t st) {
    bn_merge_kernel<<<(Cp * 32 + 127) / 128, 128, 0, st>>>(partials, n_partials, Cp, merged);
    return check_launch();
}

int bn_finalize(const float* parts, int n_parts, int C, int Cp, const float* gamma, const float* beta,
                float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                float* mean, float* invstd, float* scale, float* shift, cudaStream_t st, long long part_stride) {
    bn_finalize_kernel<<<(Cp * 32 + 127) / 128, 128, 0, st>>>(parts, n_parts, C, Cp, gamma, beta, running_mean, running_var,
                                                        num_batches_tracked, momentum, eps, mean, invstd, scale, shift,
                                                        part_stride);
    return check_launch();
}

int bn_eval_coeffs(int C, int Cp, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float eps, float* scale, float* shift, cudaStream_t st) {
    bn_eval_coeffs_kernel<<<(Cp + 127) / 128, 128, 0, st>>>(C, Cp, gamma, beta, running_mean, running_var, eps, scale, shift);
    return check_launch();
}

int bn_apply(const void* y, const void* res, void* out, int dtype, long long P, int Cp, int ld_y, int ld_res, int ld_out,
             const float* scale, const float* shift, int relu, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_y % V || ld_out % V || (res && ld_res % V)) return VAE2_ERR_ARG;
    if (Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = elemt_grid(P, Cp, V);
    if (dtype == VAE2_DT_F32)
        bn_apply_kernel<float><<<grid, BN_THREADS, kPipeBytes, st>>>((const float*)y, (const float*)res, (float*)out, P, Cp, ld_y, ld_res, ld_out, scale, shift, relu);
    else
        bn_apply_kernel<__nv_bfloat16><<<grid, BN_THREADS, kPipeBytes, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, P, Cp, ld_y, ld_res, ld_out, scale, shift, relu);
    return check_launch();
}

int bn_bwd_reduce(const void* g, const void* a, const void* y, float* partials, int* n_partials_out, int dtype,
                  long long P, int Cp, int ld_g, int ld_a, int ld_y, const float* mean, const float* invstd, int relu,
                  cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = reduce_grid(P, Cp, V);
    if (n_partials_out) *n_partials_out = grid;
    const size_t smem = kPipeBytes;   // streaming slots; the merge scratch aliases them
    if (dtype == VAE2_DT_F32)
        bn_bwd_reduce_kernel<float><<<grid, BN_THREADS, smem, st>>>((const float*)g, (const float*)a, (const float*)y, partials, P, Cp, ld_g, ld_a, ld_y, mean, invstd, relu);
    else
        bn_bwd_reduce_kernel<__nv_bfloat16><<<grid, BN_THREADS, smem, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)a, (const __nv_bfloat16*)y, partials, P, Cp, ld_g, ld_a, ld_y, mean, invstd, relu);
    return check_launch();
}

int bn_bwd_finalize(const float* partials, int n_partials, int C, int Cp, float* sums, cudaStream_t st) {
    bn_bwd_finalize_kernel<<<(Cp * 32 + 127) / 128, 128, 0, st>>>(partials, n_partials, C, Cp, sums);
    return check_launch();
}

int bn_bwd_coeffs(const float* sums, int C, int Cp, float inv_count, float* dgamma, float* dbeta, int accumulate_param,
                  const float* sums_for_param, float* c1, float* c2, cudaStream_t st) {
    bn_bwd_coeffs_kernel<<<(Cp + 127) / 128, 128, 0, st>>>(sums, C, Cp, inv_count, dgamma, dbeta, accumulate_param,
                                                          sums_for_param ? sums_for_param : sums, c1, c2);
    return check_launch();
}

int bn_bwd_elemt(const void* g, const void* a, const void* y, void* dy, void* dres, int dtype, long long P, int Cp,
                 int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean, const float* invstd,
                 const float* scale, const float* c1, const float* c2, int relu, int acc_dy, int acc_dres,
                 cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || ld_dy % V) return VAE2_ERR_ARG;
    if (Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = elemt_grid(P, Cp, V);
    if (dtype == VAE2_DT_F32)
        bn_bwd_elemt_kernel<float><<<grid, BN_THREADS, kPipeBytes, st>>>((const float*)g, (const float*)a, (const float*)y, (float*)dy, (float*)dres, P, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2, relu, acc_dy, acc_dres);
    else
        bn_bwd_elemt_kernel<__nv_bfloat16><<<grid, BN_THREADS, kPipeBytes, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)a, (const __nv_bfloat16*)y, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, P, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2, relu, acc_dy, acc_dres);
    return check_launch();
}

}  // namespace vae2
