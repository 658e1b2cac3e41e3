// Multi-resolution branch fusion (reference HighResolutionModule.forward,
// lib/models/enc_hrnet.py:233-248) and the head up-sampling (enc_hrnet.py:833-839):
//
//   out = [relu]( sum_j  resize_j(src_j) )      resize = identity or bilinear, align_corners=False
//
// One pass over the output: every source is read once (low-resolution sources through L1/L2),
// the sum is formed in registers and written once -- instead of the reference's
// interpolate + add + add + relu chain of full-resolution round trips.  With one source and
// relu=0 the same kernel is the plain bilinear up-sampler that writes a channel slice of the
// head's concat buffer.
//
// Backward is split by source kind: same-resolution sources receive g*[out>0] (one fused pass
// for all of them); up-sampled sources use a GATHER form of the transposed interpolation (each
// low-resolution pixel sums its footprint), so no atomics and deterministic results.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

constexpr int MAX_SRC = 4;

struct FuseArgs {
    const void* ptr[MAX_SRC];
    int H[MAX_SRC], W[MAX_SRC], ld[MAX_SRC];
    float sh[MAX_SRC], sw[MAX_SRC];   // input/output size ratios
    int n;
};

template <typename T>
__global__ void __launch_bounds__(256)
fuse_sum_kernel(FuseArgs a, T* __restrict__ out, int B, int H, int W, int Cp, int ld_out, int relu) {
    // one image row per blockIdx.y step, (pixel, lane vector) pairs along x: the per-element index math is 32-bit and
    // the row's vertical interpolation coefficients are computed once per thread and row (ncu r2a: the flat 64-bit
    // div/mod version was issue-bound at 72 % of the issue slots and 12 % of HBM)
    constexpr int V = Vec<T>::N;
    const unsigned lanes = (unsigned)(Cp / V);
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)W * lanes) return;
    const int x = (int)(idx / lanes);
    const int c0 = (int)(idx - (unsigned)x * lanes) * V;
    int x0s[MAX_SRC], x1s[MAX_SRC];
    float lxs[MAX_SRC];
#pragma unroll
    for (int j = 0; j < MAX_SRC; ++j) bilinear_src(x, a.sw[j], a.W[j], x0s[j], x1s[j], lxs[j]);
    for (int row = blockIdx.y; row < B * H; row += gridDim.y) {
        const int b = row / H, y = row - b * H;
        const long long p = (long long)row * W + x;
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < MAX_SRC; ++j) {
            if (j >= a.n) break;
            const T* s = reinterpret_cast<const T*>(a.ptr[j]);
            if (a.H[j] == H && a.W[j] == W) {
                const Vec<T> v = Vec<T>::load(s + p * a.ld[j] + c0);
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] += v.v[k];
            } else {
                int y0, y1;
                float ly;
                bilinear_src(y, a.sh[j], a.H[j], y0, y1, ly);
                const int x0 = x0s[j], x1 = x1s[j];
                const float lx = lxs[j];
                const long long base = (long long)b * a.H[j];
                const Vec<T> v00 = Vec<T>::load(s + ((base + y0) * a.W[j] + x0) * a.ld[j] + c0);
                const Vec<T> v01 = Vec<T>::load(s + ((base + y0) * a.W[j] + x1) * a.ld[j] + c0);
                const Vec<T> v10 = Vec<T>::load(s + ((base + y1) * a.W[j] + x0) * a.ld[j] + c0);
                const Vec<T> v11 = Vec<T>::load(s + ((base + y1) * a.W[j] + x1) * a.ld[j] + c0);
                const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
                for (int k = 0; k < V; ++k)
                    acc[k] += hy * (hx * v00.v[k] + lx * v01.v[k]) + ly * (hx * v10.v[k] + lx * v11.v[k]);
            }
        }
        Vec<T> o;
#pragma unroll
        for (int k = 0; k < V; ++k) o.v[k] = relu ? fmaxf(acc[k], 0.f) : acc[k];
        o.store(out + p * ld_out + c0);
    }
}

struct FuseDstArgs {
    void* ptr[MAX_SRC];
    int ld[MAX_SRC], acc[MAX_SRC];
    int n;
};

// same-resolution sources:  dst_j (=|+=) g * [out > 0]
template <typename T>
__global__ void __launch_bounds__(256)
fuse_bwd_same_kernel(const T* __restrict__ g, const T* __restrict__ out, FuseDstArgs d, long long P, int Cp, int ld_g,
                     int ld_out, int relu) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const long long total = P * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / lanes;
        const int c0 = (int)(i - p * lanes) * V;
        Vec<T> gv = Vec<T>::load(g + p * ld_g + c0);
        if (relu) {
            const Vec<T> ov = Vec<T>::load(out + p * ld_out + c0);
#pragma unroll
            for (int k = 0; k < V; ++k) gv.v[k] = ov.v[k] > 0.f ? gv.v[k] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < MAX_SRC; ++j) {
            if (j >= d.n) break;
            T* dp = reinterpret_cast<T*>(d.ptr[j]) + p * d.ld[j] + c0;
            Vec<T> w = gv;
            if (d.acc[j]) {
                const Vec<T> old = Vec<T>::load(dp);
#pragma unroll
                for (int k = 0; k < V; ++k) w.v[k] += old.v[k];
            }
            w.store(dp);
        }
    }
}

// up-sampled source (Hs x Ws) of an (H x W) output: gather the transposed bilinear footprint.
template <typename T>
__global__ void __launch_bounds__(256)
fuse_bwd_up_kernel(const T* __restrict__ g, const T* __restrict__ out, T* __restrict__ gsrc, int B, int H, int W,
                   int Hs, int Ws, int Cp, int ld_g, int ld_out, int ld_gsrc, float sh, float sw, int relu,
                   int accumulate) {
    constexpr int V = Vec<T>::N;
    const unsigned lanes = (unsigned)(Cp / V);
    const float rh = 1.f / sh, rw = 1.f / sw;   // output pixels per source pixel
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)Ws * lanes) return;
    const int xs = (int)(idx / lanes);
    const int c0 = (int)(idx - (unsigned)xs * lanes) * V;
    int ox_lo = (int)floorf((xs - 0.5f) * rw - 0.5f) - 1, ox_hi = (int)ceilf((xs + 1.5f) * rw - 0.5f) + 1;
    ox_lo = max(ox_lo, 0); ox_hi = min(ox_hi, W - 1);
    // trim the conservative column range to the columns that really touch xs (their weights do not depend on the row)
    while (ox_lo <= ox_hi) { int x0, x1; float lx; bilinear_src(ox_lo, sw, Ws, x0, x1, lx); if (x0 == xs || x1 == xs) break; ++ox_lo; }
    while (ox_hi >= ox_lo) { int x0, x1; float lx; bilinear_src(ox_hi, sw, Ws, x0, x1, lx); if (x0 == xs || x1 == xs) break; --ox_hi; }
    for (int row = blockIdx.y; row < B * Hs; row += gridDim.y) {
        const int b = row / Hs, ys = row - b * Hs;
        const long long p = (long long)row * Ws + xs;
        // candidate output rows whose source coordinate can fall in (ys-1, ys+1)
        int oy_lo = (int)floorf((ys - 0.5f) * rh - 0.5f) - 1, oy_hi = (int)ceilf((ys + 1.5f) * rh - 0.5f) + 1;
        oy_lo = max(oy_lo, 0); oy_hi = min(oy_hi, H - 1);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            int y0, y1; float ly;
            bilinear_src(oy, sh, Hs, y0, y1, ly);
            float wy = 0.f;
            if (y0 == ys) wy += 1.f - ly;
            if (y1 == ys) wy += ly;
            if (wy == 0.f) continue;
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                int x0, x1; float lx;
                bilinear_src(ox, sw, Ws, x0, x1, lx);
                float wx = 0.f;
                if (x0 == xs) wx += 1.f - lx;
                if (x1 == xs) wx += lx;
                if (wx == 0.f) continue;
                const long long q = ((long long)b * H + oy) * W + ox;
                Vec<T> gv = Vec<T>::load(g + q * ld_g + c0);
                if (relu) {
                    const Vec<T> ov = Vec<T>::load(out + q * ld_out + c0);
#pragma unroll
                    for (int k = 0; k < V; ++k) gv.v[k] = ov.v[k] > 0.f ? gv.v[k] : 0.f;
                }
                const float w = wy * wx;
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] = fmaf(w, gv.v[k], acc[k]);
            }
        }
        T* dp = gsrc + p * ld_gsrc + c0;
        Vec<T> o;
        if (accumulate) {
            const Vec<T> old = Vec<T>::load(dp);
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] = old.v[k] + acc[k];
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) o.v[k] = acc[k];
        }
        o.store(dp);
    }
}

int fuse_sum(const FuseSrc* srcs, int nsrc, void* out, int dtype, int B, int H, int W, int Cp, int ld_out, int relu,
             cudaStream_t st) {
    if (nsrc < 1 || nsrc > MAX_SRC) return VAE2_ERR_ARG;
    const int V = dtype == VAE2_DT_F32 ? 4 : 8;
    if (Cp % V || ld_out % V) return VAE2_ERR_ARG;
    FuseArgs a;
    a.n = nsrc;
    for (int j = 0; j < MAX_SRC; ++j) {
        const int k = j < nsrc ? j : 0;
        a.ptr[j] = srcs[k].ptr; a.H[j] = srcs[k].H; a.W[j] = srcs[k].W; a.ld[j] = srcs[k].ld;
        a.sh[j] = (float)srcs[k].H / (float)H;
        a.sw[j] = (float)srcs[k].W / (float)W;
        if (srcs[k].ld % V) return VAE2_ERR_ARG;
    }
    const int per_row = W * (Cp / V);
    const int threads = per_row >= 256 ? 256 : ((per_row + 31) / 32 * 32);
    const long long rows = (long long)B * H;
    const int gx = (per_row + threads - 1) / threads;
    long long gy = (16LL * kNumSMs * 256 / threads + gx - 1) / gx;      // ~16 CTAs per SM; threads walk rows gy apart
    if (gy > rows) gy = rows;
    dim3 grid(gx, (unsigned)(gy < 1 ? 1 : gy));
    if (dtype == VAE2_DT_F32)
        fuse_sum_kernel<float><<<grid, threads, 0, st>>>(a, (float*)out, B, H, W, Cp, ld_out, relu);
    else
        fuse_sum_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(a, (__nv_bfloat16*)out, B, H, W, Cp, ld_out, relu);
    return check_launch();
}

int fuse_bwd_same(const void* g, const void* out, const FuseDst* dsts, int ndst, int dtype, long long P, int Cp,
                  int ld_g, int ld_out, int relu, cudaStream_t st) {
    if (ndst < 1 || ndst > MAX_SRC) return VAE2_ERR_ARG;
    const int V = dtype == VAE2_DT_F32 ? 4 : 8;
    if (Cp % V || ld_g % V) return VAE2_ERR_ARG;
    FuseDstArgs d;
    d.n = ndst;
    for (int j = 0; j < MAX_SRC; ++j) {
        const int k = j < ndst ? j : 0;
        d.ptr[j] = dsts[k].ptr; d.ld[j] = dsts[k].ld; d.acc[j] = dsts[k].accumulate;
    }
    const int grid = stream_grid(P * (Cp / V), 256 * 2);
    if (dtype == VAE2_DT_F32)
        fuse_bwd_same_kernel<float><<<grid, 256, 0, st>>>((const float*)g, (const float*)out, d, P, Cp, ld_g, ld_out, relu);
    else
        fuse_bwd_same_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)out, d, P, Cp, ld_g, ld_out, relu);
    return check_launch();
}

int fuse_bwd_up(const void* g, const void* out, void* gsrc, int dtype, int B, int H, int W, int Hs, int Ws, int Cp,
                int ld_g, int ld_out, int ld_gsrc, int relu, int accumulate, cudaStream_t st) {
    const int V = dtype == VAE2_DT_F32 ? 4 : 8;
    if (Cp % V || ld_g % V || ld_gsrc % V) return VAE2_ERR_ARG;
    const float sh = (float)Hs / (float)H, sw = (float)Ws / (float)W;
    const int per_row = Ws * (Cp / V);
    const int threads = per_row >= 256 ? 256 : ((per_row + 31) / 32 * 32);
    const long long rows = (long long)B * Hs;
    const int gx = (per_row + threads - 1) / threads;
    long long gy = (16LL * kNumSMs * 256 / threads + gx - 1) / gx;
    if (gy > rows) gy = rows;
    dim3 grid(gx, (unsigned)(gy < 1 ? 1 : gy));
    if (dtype == VAE2_DT_F32)
        fuse_bwd_up_kernel<float><<<grid, threads, 0, st>>>((const float*)g, (const float*)out, (float*)gsrc, B, H, W, Hs, Ws, Cp, ld_g, ld_out, ld_gsrc, sh, sw, relu, accumulate);
    else
        fuse_bwd_up_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)out, (__nv_bfloat16*)gsrc, B, H, W, Hs, Ws, Cp, ld_g, ld_out, ld_gsrc, sh, sw, relu, accumulate);
    return check_launch();
}

}  // namespace vae2
