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
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_PARTS = kNumSMs * 2;   // <= 2 CTAs per SM: few, fat partials keep the merges short

int bn_stats_max_partials() { return BN_MAX_PARTS; }

// ---------------------------------------------------------------------------
// statistics: per-CTA (count, mean, M2) per channel  -> partials[cta][3][Cp]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const T* __restrict__ y, float* __restrict__ partials, long long P, int Cp, int ld) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const int R = BN_THREADS / lanes;
    const int lane = threadIdx.x % lanes, r = threadIdx.x / lanes;
    const bool active = r < R;

    float mean[V], m2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { mean[i] = 0.f; m2[i] = 0.f; }
    float n = 0.f;
    if (active) {
        // 4 independent 16-byte loads in flight per thread before the (serial) Welford updates
        const long long stride = (long long)gridDim.x * R;
        for (long long p = (long long)blockIdx.x * R + r; p < P; p += 4 * stride) {
            Vec<T> x[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long q = p + u * stride;
                ok[u] = q < P;
                if (ok[u]) x[u] = Vec<T>::load(y + q * ld + lane * V);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                n += 1.f;
                const float inv = 1.f / n;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float d = x[u].v[i] - mean[i];
                    mean[i] += d * inv;
                    m2[i] = fmaf(d, x[u].v[i] - mean[i], m2[i]);
                }
            }
        }
    }
    // CTA merge (Chan): thread rows r = 0..R-1 for each lane, folded by row 0
    extern __shared__ float sm[];          // [BN_THREADS][2*V + 1]
    float* mine = sm + threadIdx.x * (2 * V + 1);
#pragma unroll
    for (int i = 0; i < V; ++i) { mine[i] = mean[i]; mine[V + i] = m2[i]; }
    mine[2 * V] = n;
    __syncthreads();
    if (active && r == 0) {
        for (int rr = 1; rr < R; ++rr) {
            const float* o = sm + (rr * lanes + lane) * (2 * V + 1);
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
            }
        }
        float* out = partials + (long long)blockIdx.x * 3 * Cp;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = lane * V + i;
            out[c] = n;
            out[Cp + c] = mean[i];
            out[2 * Cp + c] = m2[i];
        }
    }
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

// One WARP per channel: lanes fold partials lane, lane+32, ... then a 5-step shuffle tree of Chan
// merges (the merge is associative), so the latency is ~n_parts/32 dependent loads, not n_parts.
__device__ __forceinline__ void warp_merge_parts(const float* __restrict__ parts, int n_parts, int Cp, int c,
                                                 float& n, float& mean, float& m2, long long part_stride = 0) {
    if (part_stride == 0) part_stride = 3LL * Cp;
    const int lane = threadIdx.x & 31;
    n = 0.f; mean = 0.f; m2 = 0.f;
    for (int k = lane; k < n_parts; k += 64) {   // two independent partials in flight per step
        const float* p = parts + (long long)k * part_stride;
        const bool has2 = k + 32 < n_parts;
        const float* q = parts + (long long)(has2 ? k + 32 : k) * part_stride;
        const float a0 = p[c], a1 = p[Cp + c], a2 = p[2 * Cp + c];
        const float b0 = q[c], b1 = q[Cp + c], b2 = q[2 * Cp + c];
        chan_merge(n, mean, m2, a0, a1, a2);
        if (has2) chan_merge(n, mean, m2, b0, b1, b2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o);
        const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
        const float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        // symmetric form so that both partners end with the same value
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

// merge n_parts partial sets -> one (count, mean, M2) set  (per-rank partial for SyncBN all-gather)
__global__ void bn_merge_kernel(const float* __restrict__ parts, int n_parts, int Cp, float* __restrict__ merged) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= Cp) return;
    float n, mean, m2;
    warp_merge_parts(parts, n_parts, Cp, c, n, mean, m2);
    if ((threadIdx.x & 31) == 0) { merged[c] = n; merged[Cp + c] = mean; merged[2 * Cp + c] = m2; }
}

// merge + produce normalisation coefficients + running-stat update
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
    const bool lead = (threadIdx.x & 31) == 0;
    if (c >= C) {
        if (lead) { mean_o[c] = 0.f; invstd_o[c] = 0.f; scale_o[c] = 0.f; shift_o[c] = 0.f; }
        return;
    }
    float n, mean, m2;
    warp_merge_parts(parts, n_parts, Cp, c, n, mean, m2, part_stride);
    if (!lead) return;
    const float var = m2 / n;
    const float invstd = rsqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    mean_o[c] = mean; invstd_o[c] = invstd; scale_o[c] = sc; shift_o[c] = beta[c] - mean * sc;
    if (running_mean != nullptr) {
        const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
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
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const T* __restrict__ y, const T* __restrict__ res, T* __restrict__ out, long long P, int Cp,
                int ld_y, int ld_res, int ld_out, const float* __restrict__ scale, const float* __restrict__ shift,
                int relu) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const long long total = P * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / lanes;
        const int c0 = (int)(i - p * lanes) * V;
        Vec<T> x = Vec<T>::load(y + p * ld_y + c0);
        Vec<T> o;
#pragma unroll
        for (int k = 0; k < V; ++k) o.v[k] = fmaf(x.v[k], __ldg(scale + c0 + k), __ldg(shift + c0 + k));
        if (res != nullptr) {
            const Vec<T> rr = Vec<T>::load(res + p * ld_res + c0);
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] += rr.v[k];
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] = fmaxf(o.v[k], 0.f);
        }
        o.store(out + p * ld_out + c0);
    }
}

// ---------------------------------------------------------------------------
// backward pass 1: per-CTA  sum(dyb), sum(dyb * xhat)   dyb = g * [a > 0]
// partials[cta][2][Cp]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y,
                     float* __restrict__ partials, long long P, int Cp, int ld_g, int ld_a, int ld_y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const int R = BN_THREADS / lanes;
    const int lane = threadIdx.x % lanes, r = threadIdx.x / lanes;
    const bool active = r < R;
    float s1[V], s2[V], mu[V], is[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    if (active) {
#pragma unroll
        for (int i = 0; i < V; ++i) { mu[i] = mean[lane * V + i]; is[i] = invstd[lane * V + i]; }
        const long long stride = (long long)gridDim.x * R;
        for (long long p = (long long)blockIdx.x * R + r; p < P; p += 2 * stride) {
            Vec<T> gv[2], yv[2], av[2];
            bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {     // 2 pixels x 3 tensors = 6 independent loads in flight
                const long long q = p + u * stride;
                ok[u] = q < P;
                if (ok[u]) {
                    gv[u] = Vec<T>::load(g + q * ld_g + lane * V);
                    yv[u] = Vec<T>::load(y + q * ld_y + lane * V);
                    if (relu) av[u] = Vec<T>::load(a + q * ld_a + lane * V);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!ok[u]) continue;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float gg = (relu && !(av[u].v[i] > 0.f)) ? 0.f : gv[u].v[i];
                    s1[i] += gg;
                    s2[i] = fmaf(gg, (yv[u].v[i] - mu[i]) * is[i], s2[i]);
                }
            }
        }
    }
    extern __shared__ float sm[];  // [BN_THREADS][2*V]
    float* mine = sm + threadIdx.x * (2 * V);
#pragma unroll
    for (int i = 0; i < V; ++i) { mine[i] = s1[i]; mine[V + i] = s2[i]; }
    __syncthreads();
    if (active && r == 0) {
        for (int rr = 1; rr < R; ++rr) {
            const float* o = sm + (rr * lanes + lane) * (2 * V);
#pragma unroll
            for (int i = 0; i < V; ++i) { s1[i] += o[i]; s2[i] += o[V + i]; }
        }
        float* out = partials + (long long)blockIdx.x * 2 * Cp;
#pragma unroll
        for (int i = 0; i < V; ++i) { out[lane * V + i] = s1[i]; out[Cp + lane * V + i] = s2[i]; }
    }
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
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_elemt_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y, T* __restrict__ dy,
                    T* __restrict__ dres, long long P, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
                    const float* __restrict__ c1, const float* __restrict__ c2, int relu, int acc_dy, int acc_dres) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const long long total = P * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / lanes;
        const int c0 = (int)(i - p * lanes) * V;
        Vec<T> gv = Vec<T>::load(g + p * ld_g + c0);
        const Vec<T> yv = Vec<T>::load(y + p * ld_y + c0);
        if (relu) {
            const Vec<T> av = Vec<T>::load(a + p * ld_a + c0);
#pragma unroll
            for (int k = 0; k < V; ++k) gv.v[k] = av.v[k] > 0.f ? gv.v[k] : 0.f;
        }
        Vec<T> o;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = c0 + k;
            const float xh = (yv.v[k] - __ldg(mean + c)) * __ldg(invstd + c);
            o.v[k] = __ldg(scale + c) * (gv.v[k] - __ldg(c1 + c) - xh * __ldg(c2 + c));
        }
        if (acc_dy) {
            const Vec<T> old = Vec<T>::load(dy + p * ld_dy + c0);
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] += old.v[k];
        }
        o.store(dy + p * ld_dy + c0);
        if (dres != nullptr) {
            if (acc_dres) {
                const Vec<T> old = Vec<T>::load(dres + p * ld_dres + c0);
#pragma unroll
                for (int k = 0; k < V; ++k) gv.v[k] += old.v[k];
            }
            gv.store(dres + p * ld_dres + c0);
        }
    }
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
    if (need > BN_MAX_PARTS) need = BN_MAX_PARTS;
    return (int)need;
}

int bn_stats(const void* y, float* partials, int* n_partials_out, int dtype, long long P, int Cp, int ld, cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld % V || Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = reduce_grid(P, Cp, V);
    if (n_partials_out) *n_partials_out = grid;
    const size_t smem = (size_t)BN_THREADS * (2 * V + 1) * sizeof(float);
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
    const int grid = stream_grid(P * (Cp / V), BN_THREADS * 4);
    if (dtype == VAE2_DT_F32)
        bn_apply_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)y, (const float*)res, (float*)out, P, Cp, ld_y, ld_res, ld_out, scale, shift, relu);
    else
        bn_apply_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, P, Cp, ld_y, ld_res, ld_out, scale, shift, relu);
    return check_launch();
}

int bn_bwd_reduce(const void* g, const void* a, const void* y, float* partials, int* n_partials_out, int dtype,
                  long long P, int Cp, int ld_g, int ld_a, int ld_y, const float* mean, const float* invstd, int relu,
                  cudaStream_t st) {
    const int V = vec_of(dtype);
    if (Cp % V || ld_g % V || ld_y % V || Cp / V > BN_THREADS) return VAE2_ERR_ARG;
    const int grid = reduce_grid(P, Cp, V);
    if (n_partials_out) *n_partials_out = grid;
    const size_t smem = (size_t)BN_THREADS * 2 * V * sizeof(float);
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
    const int grid = stream_grid(P * (Cp / V), BN_THREADS * 4);
    if (dtype == VAE2_DT_F32)
        bn_bwd_elemt_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)g, (const float*)a, (const float*)y, (float*)dy, (float*)dres, P, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2, relu, acc_dy, acc_dres);
    else
        bn_bwd_elemt_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)a, (const __nv_bfloat16*)y, (__nv_bfloat16*)dy, (__nv_bfloat16*)dres, P, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2, relu, acc_dy, acc_dres);
    return check_launch();
}

}  // namespace vae2
