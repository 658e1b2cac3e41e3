// Shared device helpers for the VAE^2 B200 kernels (sm_100a).
//
// Activation layout everywhere in this library: channels-last, [B][H][W][ld] with
// `ld` >= Cp (padded channel count) elements between consecutive pixels, so a
// tensor can live inside a channel slice of a wider (concat) buffer.  Pad lanes
// (c >= C) always hold zero.  Storage type is fp32 (dtype 0) or bf16 (dtype 1);
// all arithmetic is fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define VAE2_OK 0
#define VAE2_ERR_ARG 1
#define VAE2_ERR_CUDA 2
#define VAE2_ERR_UNSUPPORTED 3

#define VAE2_DT_F32 0
#define VAE2_DT_BF16 1

namespace vae2 {

constexpr int kNumSMs = 148;  // B200

// ---- vector-of-channels access: VEC channels per thread, 16-byte transactions ----
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ static Vec load(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ static Vec load(const __nv_bfloat16* p) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
        Vec r;
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
        return r;
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};

// Raw 16-byte vectors: streaming kernels keep loads packed in registers and unpack element k (a compile-time
// index after unrolling) at the point of use, so many loads can be in flight without 2x the registers.
template <typename T> __device__ __forceinline__ uint4 ld16(const T* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ unsigned word_of(const uint4& q, int w) { return w == 0 ? q.x : w == 1 ? q.y : w == 2 ? q.z : q.w; }
template <typename T> __device__ __forceinline__ float elem(const uint4& q, int k);
template <> __device__ __forceinline__ float elem<float>(const uint4& q, int k) { return __uint_as_float(word_of(q, k)); }
template <> __device__ __forceinline__ float elem<__nv_bfloat16>(const uint4& q, int k) {
    const unsigned w = word_of(q, k >> 1);
    return __uint_as_float((k & 1) ? (w & 0xffff0000u) : (w << 16));
}

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of one float per thread; result valid in thread 0. `red` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// Division by a run-time constant without the ~25-instruction integer divide: q = n / d for 0 <= n < 2^31
// (Granlund-Montgomery round-up method: m = floor(2^32 * (2^l - d) / d) + 1, q = (mulhi(m, n) + n) >> l).
struct FastDiv {
    unsigned d, m, l;
};
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    unsigned l = 0;
    while ((1ull << l) < d) ++l;
    f.l = l;
    f.m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ unsigned fast_div(unsigned n, const FastDiv& f) {
    return (unsigned)(((unsigned long long)__umulhi(f.m, n) + n) >> f.l);
}

// Grid sizing for streaming kernels: a multiple of the SM count, capped by the work.
inline int stream_grid(long long work_items, int per_block, int waves = 8) {
    long long need = (work_items + per_block - 1) / per_block;
    long long cap = (long long)kNumSMs * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Name of the kernel the last launcher on this thread picked (introspection for the in-step profiler:
// vae2_last_kernel() in the C ABI).  Launch sites with several candidate kernels note their choice.
extern thread_local const char* g_last_kernel;
inline void note_kernel(const char* name) { g_last_kernel = name; }

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VAE2_OK : VAE2_ERR_CUDA;
}

// PyTorch's bilinear source index (align_corners=False): area_pixel_compute_source_index.
__device__ __forceinline__ void bilinear_src(int o, float scale, int in_size, int& i0, int& i1, float& lam) {
    float s = scale * (o + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    lam = s - (float)i0;
}

}  // namespace vae2
