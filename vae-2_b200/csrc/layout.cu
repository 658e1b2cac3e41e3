// Layout kernels: NCHW fp32 (the reference's boundary layout) <-> channels-last padded
// internal tensors, channel-slice copies, per-sample code broadcast, and the batched
// weight pack / weight-gradient unpack that maps the reference's OIHW parameters
// (lib/models/enc_hrnet.py:27-30 nn.Conv2d weights) onto GEMM operand layouts.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

// ---------------------------------------------------------------------------
// NCHW fp32 -> NHWC(T).  One thread per pixel; reads are coalesced across the warp
// for each channel plane, writes are Cp contiguous elements per thread.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst,
                                    int B, int C, int Cp, int HW, int ld, int src_ctot, int src_coff) {
    const long long total = (long long)B * HW;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const int hw = (int)(p - (long long)b * HW);
        const float* s = src + ((long long)b * src_ctot + src_coff) * HW + hw;
        T* d = dst + p * ld;
        for (int c = 0; c < Cp; ++c) d[c] = from_f<T>(c < C ? s[(long long)c * HW] : 0.f);
    }
}

// NHWC(T) -> NCHW fp32 (dst channel window [dst_coff, dst_coff+C) of dst_ctot), optional accumulate.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                    int B, int C, int HW, int ld, int dst_ctot, int dst_coff, int accumulate) {
    const long long total = (long long)B * HW;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const int hw = (int)(p - (long long)b * HW);
        const T* s = src + p * ld;
        float* d = dst + ((long long)b * dst_ctot + dst_coff) * HW + hw;
        for (int c = 0; c < C; ++c) {
            float v = to_f<T>(s[c]);
            if (accumulate) v += d[(long long)c * HW];
            d[(long long)c * HW] = v;
        }
    }
}

// dst[p][0..Cp) (=|+=) src[p][0..Cp) with independent pixel strides (channel-slice copy).
template <typename T>
__global__ void slice_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, long long P, int Cp,
                                  int ld_src, int ld_dst, int accumulate) {
    constexpr int V = Vec<T>::N;
    const int lanes = Cp / V;
    const long long total = P * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / lanes;
        const int l = (int)(i - p * lanes);
        Vec<T> a = Vec<T>::load(src + p * ld_src + l * V);
        if (accumulate) {
            Vec<T> b = Vec<T>::load(dst + p * ld_dst + l * V);
#pragma unroll
            for (int k = 0; k < V; ++k) a.v[k] += b.v[k];
        }
        a.store(dst + p * ld_dst + l * V);
    }
}

// Code map (reference _gen_code_map, enc_hrnet.py:454-462): dst[b][hw][0..Z) = code[b][0..Z).
template <typename T>
__global__ void code_broadcast_kernel(const float* __restrict__ code, T* __restrict__ dst, int B, int Z, int Zp,
                                      int HW, int ld) {
    const long long total = (long long)B * HW;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        T* d = dst + p * ld;
        for (int c = 0; c < Zp; ++c) d[c] = from_f<T>(c < Z ? code[b * Z + c] : 0.f);
    }
}

// ---------------------------------------------------------------------------
// Batched weight pack.  For conv i (grid.y): OIHW fp32 -> wp[tap][Cin_p][Cout_p] (forward /
// wgrad operand) and wpT[tap][Cout_p][Cin_p] (dgrad operand), input channel c going to
// physical lane cin_map[c] (identity when null).  Pad lanes stay zero (buffers are
// zero-filled once at plan build).  Optional bf16 copies for the tensor-core path:
// wq[tap][Cout_p][Cin_p] (K-major B operand of the forward implicit GEMM).
// ---------------------------------------------------------------------------
__global__ void pack_weights_kernel(const PackDesc* __restrict__ descs) {
    const PackDesc d = descs[blockIdx.y];
    const int kk = d.k * d.k;
    const int total = d.Cout * d.Cin * kk;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % kk;
        const int ci = (i / kk) % d.Cin;
        const int co = i / (kk * d.Cin);
        const float w = d.w[i];
        const int pc = d.cin_map ? d.cin_map[ci] : ci;
        if (d.wp) d.wp[((long long)tap * d.Cin_p + pc) * d.Cout_p + co] = w;
        if (d.wpT) d.wpT[((long long)tap * d.Cout_p + co) * d.Cin_p + pc] = w;
        if (d.wq) d.wq[((long long)tap * d.Cout_p + co) * d.Cin_p + pc] = __float2bfloat16_rn(w);
        if (d.wqT) d.wqT[((long long)tap * d.Cin_p + pc) * d.Cout_p + co] = __float2bfloat16_rn(w);
    }
}

// Batched inverse for weight gradients: dwp[tap][Cin_p][Cout_p] -> dW OIHW (=|+=).
__global__ void unpack_wgrad_kernel(const PackDesc* __restrict__ descs, int accumulate) {
    const PackDesc d = descs[blockIdx.y];
    const int kk = d.k * d.k;
    const int total = d.Cout * d.Cin * kk;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % kk;
        const int ci = (i / kk) % d.Cin;
        const int co = i / (kk * d.Cin);
        const int pc = d.cin_map ? d.cin_map[ci] : ci;
        float g = d.wp[((long long)tap * d.Cin_p + pc) * d.Cout_p + co];
        float* dst = const_cast<float*>(d.w) + i;
        *dst = accumulate ? (*dst + g) : g;
    }
}

template <typename T>
static int launch_nchw_to_nhwc(const float* src, void* dst, int B, int C, int Cp, int H, int W, int ld,
                               int src_ctot, int src_coff, cudaStream_t st) {
    const long long P = (long long)B * H * W;
    nchw_to_nhwc_kernel<T><<<stream_grid(P, 256), 256, 0, st>>>(src, (T*)dst, B, C, Cp, H * W, ld, src_ctot, src_coff);
    return check_launch();
}

int nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int Cp, int H, int W, int ld,
                 int src_ctot, int src_coff, cudaStream_t st) {
    if (dtype == VAE2_DT_F32) return launch_nchw_to_nhwc<float>(src, dst, B, C, Cp, H, W, ld, src_ctot, src_coff, st);
    return launch_nchw_to_nhwc<__nv_bfloat16>(src, dst, B, C, Cp, H, W, ld, src_ctot, src_coff, st);
}

int nhwc_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W, int ld, int dst_ctot,
                 int dst_coff, int accumulate, cudaStream_t st) {
    const long long P = (long long)B * H * W;
    const int g = stream_grid(P, 256);
    if (dtype == VAE2_DT_F32)
        nhwc_to_nchw_kernel<float><<<g, 256, 0, st>>>((const float*)src, dst, B, C, H * W, ld, dst_ctot, dst_coff, accumulate);
    else
        nhwc_to_nchw_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, dst, B, C, H * W, ld, dst_ctot, dst_coff, accumulate);
    return check_launch();
}

int slice_copy(const void* src, void* dst, int dtype, long long P, int Cp, int ld_src, int ld_dst, int accumulate,
               cudaStream_t st) {
    if (dtype == VAE2_DT_F32) {
        if (Cp % 4 || ld_src % 4 || ld_dst % 4) return VAE2_ERR_ARG;
        slice_copy_kernel<float><<<stream_grid(P * (Cp / 4), 256), 256, 0, st>>>((const float*)src, (float*)dst, P, Cp, ld_src, ld_dst, accumulate);
    } else {
        if (Cp % 8 || ld_src % 8 || ld_dst % 8) return VAE2_ERR_ARG;
        slice_copy_kernel<__nv_bfloat16><<<stream_grid(P * (Cp / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, P, Cp, ld_src, ld_dst, accumulate);
    }
    return check_launch();
}

int code_broadcast(const float* code, void* dst, int dtype, int B, int Z, int Zp, int H, int W, int ld, cudaStream_t st) {
    const long long P = (long long)B * H * W;
    const int g = stream_grid(P, 256);
    if (dtype == VAE2_DT_F32)
        code_broadcast_kernel<float><<<g, 256, 0, st>>>(code, (float*)dst, B, Z, Zp, H * W, ld);
    else
        code_broadcast_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(code, (__nv_bfloat16*)dst, B, Z, Zp, H * W, ld);
    return check_launch();
}

int pack_weights(const PackDesc* descs_dev, int n, cudaStream_t st) {
    if (n <= 0) return VAE2_OK;
    dim3 grid(8, n);
    pack_weights_kernel<<<grid, 256, 0, st>>>(descs_dev);
    return check_launch();
}

int unpack_wgrad(const PackDesc* descs_dev, int n, int accumulate, cudaStream_t st) {
    if (n <= 0) return VAE2_OK;
    dim3 grid(8, n);
    unpack_wgrad_kernel<<<grid, 256, 0, st>>>(descs_dev, accumulate);
    return check_launch();
}


// ---------------------------------------------------------------------------
// Per-sample spatial reductions / broadcasts (reference: nn.AdaptiveAvgPool2d((1,1)) of the non-HD_Z posterior head,
// enc_hrnet.py:1023-1041, and the gradient of `_gen_code_map`'s spatial repeat of a per-sample z, :454-462).
//   spatial_sum  : out[b][c] (=|+=) scale * sum_p x[b][p][c]     x act (T), out fp32 [B][out_ld] or act (T) [B][out_ld]
//   spatial_bcast: dx[b][p][c] (=|+=) scale * g[b][c]            g fp32 or act (T), dx act (T)
// ---------------------------------------------------------------------------
template <typename T, typename U>
__global__ void __launch_bounds__(256)
spatial_sum_kernel(const T* __restrict__ x, U* __restrict__ out, int HW, int C, int ld, int out_ld, float scale, int accumulate) {
    __shared__ float red[8][33];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), row = threadIdx.x >> 5;
    float s = 0.f;
    if (c < C)
        for (int p = row; p < HW; p += 8) s += to_f<T>(x[((long long)b * HW + p) * ld + c]);
    red[row][threadIdx.x & 31] = s;
    __syncthreads();
    if (row == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x & 31];
        t *= scale;
        U* o = out + (long long)b * out_ld + c;
        if (accumulate) t += to_f<U>(*o);
        *o = from_f<U>(t);
    }
}

template <typename T, typename U>
__global__ void __launch_bounds__(256)
spatial_bcast_kernel(const U* __restrict__ g, T* __restrict__ dx, int HW, int C, int Cp, int ld, int g_ld, float scale,
                     int accumulate, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long long p = i / Cp;
        const int b = (int)(p / HW);
        float v = c < C ? scale * to_f<U>(g[(long long)b * g_ld + c]) : 0.f;
        T* d = dx + p * ld + c;
        if (accumulate) v += to_f<T>(*d);
        *d = from_f<T>(v);
    }
}

int spatial_sum(const void* x, void* out, int dtype, int out_fp32, int B, int HW, int C, int ld, int out_ld, float scale,
                int accumulate, cudaStream_t st) {
    if (B < 1 || HW < 1 || C < 1) return VAE2_ERR_ARG;
    dim3 grid((C + 31) / 32, B);
    if (dtype == VAE2_DT_F32)
        spatial_sum_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, HW, C, ld, out_ld, scale, accumulate);
    else if (out_fp32)
        spatial_sum_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)out, HW, C, ld, out_ld, scale, accumulate);
    else
        spatial_sum_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, HW, C, ld, out_ld, scale, accumulate);
    return check_launch();
}

int spatial_bcast(const void* g, void* dx, int dtype, int g_fp32, int B, int HW, int C, int Cp, int ld, int g_ld, float scale,
                  int accumulate, cudaStream_t st) {
    const long long total = (long long)B * HW * Cp;
    if (total < 1) return VAE2_ERR_ARG;
    const int grid = stream_grid(total, 256);
    if (dtype == VAE2_DT_F32)
        spatial_bcast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)g, (float*)dx, HW, C, Cp, ld, g_ld, scale, accumulate, total);
    else if (g_fp32)
        spatial_bcast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const float*)g, (__nv_bfloat16*)dx, HW, C, Cp, ld, g_ld, scale, accumulate, total);
    else
        spatial_bcast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)dx, HW, C, Cp, ld, g_ld, scale, accumulate, total);
    return check_launch();
}

}  // namespace vae2
