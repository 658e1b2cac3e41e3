// Caller-side kernels either side of the path:
//   * clip_u8_to_nchw   : the dataset's frame preparation (lib/datasets/cityscapes.py:300-326: /255, ImageNet mean/std,
//                          HWC -> CHW, three frames stacked along channels) on the device, from uint8 frames;
//   * to_image          : `_to_image(x, is_uint8=False)` of the inference driver (lib/core/function.py:87-98):
//                          (x*std + mean)*255 clipped to [0, 255];
//   * frame_metrics     : per (sample, frame) mean |im - im_gt| and PSNR inputs (function.py:262-263; criterion.py:106-116);
//   * ssim_level / avgpool2 : SSIM and the levels of MS-SSIM as pytorch_msssim computes them (function.py:244-261 calls
//                          ssim / ms_ssim with data_range=255, size_average=True, ms_ssim weights [1/3]*3): 11-tap
//                          Gaussian (sigma 1.5), separable, VALID convolution per channel, K1=0.01, K2=0.03.
// All HBM-bound streaming / small-stencil work: coalesced, shared-memory staged, grids sized from the SM count.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

__constant__ float c_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_std[3] = {0.229f, 0.224f, 0.225f};

// src uint8 [B][L][H][W][3] (RGB, HWC) -> dst fp32 [B][3L][H][W]
__global__ void __launch_bounds__(256)
clip_u8_to_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long npix_total, int HW, int L) {
    // one thread per pixel of one frame: reads 3 contiguous bytes, writes 3 planes (each plane write is coalesced)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix_total; i += (long long)gridDim.x * blockDim.x) {
        const long long frame = i / HW;            // b*L + l
        const int p = (int)(i - frame * HW);
        const uint8_t* s = src + i * 3;
        float* d = dst + frame * 3 * (long long)HW + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) d[(long long)c * HW] = ((float)s[c] * (1.0f / 255.0f) - c_mean[c]) / c_std[c];
    }
}

int clip_u8_to_nchw(const uint8_t* src, float* dst, int B, int L, int H, int W, cudaStream_t st) {
    const long long n = (long long)B * L * H * W;
    if (n <= 0) return VAE2_OK;
    clip_u8_to_nchw_kernel<<<stream_grid(n, 256), 256, 0, st>>>(src, dst, n, H * W, L);
    return check_launch();
}

// x fp32 [N][C3][HW] with channel c%3 -> image space
__global__ void __launch_bounds__(256)
to_image_kernel(const float* __restrict__ x, float* __restrict__ im, long long n, int HW) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % 3);
        float v = (x[i] * c_std[c] + c_mean[c]) * 255.0f;      // same operation order as the reference's numpy code
        im[i] = fminf(fmaxf(v, 0.f), 255.f);
    }
}

int to_image(const float* x, float* im, long long n, int HW, cudaStream_t st) {
    if (n <= 0) return VAE2_OK;
    to_image_kernel<<<stream_grid(n, 256), 256, 0, st>>>(x, im, n, HW);
    return check_launch();
}

// pred image planes [R][F][3][HW], gt [Bg][F][3][HW] (row r compares with gt row r % Bg): out[(r*F+f)*2 + {0,1}] += sum|d|, sum d^2
__global__ void __launch_bounds__(256)
frame_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ out, int R, int F,
                     int Bg, int frame_elems, int chunks) {
    __shared__ float red[2][8];
    const int rf = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int r = rf / F, f = rf % F;
    const float* p = pred + (long long)rf * frame_elems;
    const float* g = gt + ((long long)(r % Bg) * F + f) * frame_elems;
    float s1 = 0.f, s2 = 0.f;
    for (int i = chunk * blockDim.x + threadIdx.x; i < frame_elems; i += chunks * blockDim.x) {
        const float d = p[i] - g[i];
        s1 += fabsf(d);
        s2 += d * d;
    }
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(~0u, s1, o); s2 += __shfl_xor_sync(~0u, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(out + (long long)rf * 2, a);
        atomicAdd(out + (long long)rf * 2 + 1, b);
    }
}

int frame_metrics(const float* pred, const float* gt, double* out, int R, int F, int Bg, int frame_elems, cudaStream_t st) {
    if (R <= 0) return VAE2_OK;
    int chunks = (2 * kNumSMs + R * F - 1) / (R * F);
    if (chunks < 1) chunks = 1;
    const int maxc = (frame_elems + 255) / 256;
    if (chunks > maxc) chunks = maxc;
    cudaMemsetAsync(out, 0, sizeof(double) * 2 * R * F, st);
    frame_metrics_kernel<<<R * F * chunks, 256, 0, st>>>(pred, gt, out, R, F, Bg, frame_elems, chunks);
    return check_launch();
}

// ---- SSIM ------------------------------------------------------------------------------------------------------------------
constexpr int kWin = 11, kTile = 16, kHalo = kTile + kWin - 1;   // 26

struct Gauss { float w[kWin]; };

// X planes [N][H][W] (N = images*channels), Y planes [Ny][H][W] (plane n compares with Y plane n % Ny).
// out[n*2 + 0] += sum of ssim_map, out[n*2 + 1] += sum of cs_map over the (H-10) x (W-10) valid positions.
__global__ void __launch_bounds__(kTile * kTile)
ssim_level_kernel(const float* __restrict__ X, const float* __restrict__ Y, double* __restrict__ out, int Ny, int H, int W,
                  int tiles_w, int tiles_h, float C1, float C2, Gauss gw) {
    __shared__ float sx[kHalo][kHalo + 1], sy[kHalo][kHalo + 1];
    __shared__ float hx[kHalo][kTile], hy[kHalo][kTile], hxx[kHalo][kTile], hyy[kHalo][kTile], hxy[kHalo][kTile];
    __shared__ float red[2][8];
    const int n = blockIdx.x / (tiles_w * tiles_h);
    const int t = blockIdx.x % (tiles_w * tiles_h);
    const int oh0 = (t / tiles_w) * kTile, ow0 = (t % tiles_w) * kTile;
    const int Ho = H - kWin + 1, Wo = W - kWin + 1;
    const float* x = X + (long long)n * H * W;
    const float* y = Y + (long long)(n % Ny) * H * W;
    const int tid = threadIdx.y * kTile + threadIdx.x;
    for (int i = tid; i < kHalo * kHalo; i += kTile * kTile) {
        const int r = i / kHalo, c = i % kHalo;
        const int h = oh0 + r, w = ow0 + c;
        const bool ok = h < H && w < W;
        sx[r][c] = ok ? x[(long long)h * W + w] : 0.f;
        sy[r][c] = ok ? y[(long long)h * W + w] : 0.f;
    }
    __syncthreads();
    // horizontal pass: kHalo rows x kTile columns
    for (int i = tid; i < kHalo * kTile; i += kTile * kTile) {
        const int r = i / kTile, c = i % kTile;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float u = sx[r][c + k], v = sy[r][c + k], g = gw.w[k];
            a += g * u; b += g * v; aa += g * u * u; bb += g * v * v; ab += g * u * v;
        }
        hx[r][c] = a; hy[r][c] = b; hxx[r][c] = aa; hyy[r][c] = bb; hxy[r][c] = ab;
    }
    __syncthreads();
    float s_ssim = 0.f, s_cs = 0.f;
    {
        const int r = threadIdx.y, c = threadIdx.x;
        if (oh0 + r < Ho && ow0 + c < Wo) {
            float mu1 = 0.f, mu2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float g = gw.w[k];
                mu1 += g * hx[r + k][c]; mu2 += g * hy[r + k][c];
                xx += g * hxx[r + k][c]; yy += g * hyy[r + k][c]; xy += g * hxy[r + k][c];
            }
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = xx - mu1_sq, s2 = yy - mu2_sq, s12 = xy - mu12;
            const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
            s_cs = cs;
            s_ssim = ((2.f * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs;
        }
    }
    for (int o = 16; o; o >>= 1) { s_ssim += __shfl_xor_sync(~0u, s_ssim, o); s_cs += __shfl_xor_sync(~0u, s_cs, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_ssim; red[1][tid >> 5] = s_cs; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(out + (long long)n * 2, a);
        atomicAdd(out + (long long)n * 2 + 1, b);
    }
}

int ssim_level(const float* X, const float* Y, double* out, int N, int Ny, int H, int W, float data_range, cudaStream_t st) {
    if (H < kWin || W < kWin) return VAE2_ERR_ARG;
    Gauss gw;
    double sum = 0;
    for (int i = 0; i < kWin; ++i) { const double c = i - kWin / 2; gw.w[i] = (float)exp(-(c * c) / (2.0 * 1.5 * 1.5)); sum += gw.w[i]; }
    for (int i = 0; i < kWin; ++i) gw.w[i] = (float)(gw.w[i] / sum);
    const int Ho = H - kWin + 1, Wo = W - kWin + 1;
    const int tw = (Wo + kTile - 1) / kTile, th = (Ho + kTile - 1) / kTile;
    const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
    cudaMemsetAsync(out, 0, sizeof(double) * 2 * N, st);
    ssim_level_kernel<<<N * tw * th, dim3(kTile, kTile), 0, st>>>(X, Y, out, Ny, H, W, tw, th, C1, C2, gw);
    return check_launch();
}

// F.avg_pool2d(x, kernel_size=2, padding=(H%2, W%2)) (count_include_pad=True): planes [N][H][W] -> [N][Ho][Wo]
__global__ void __launch_bounds__(256)
avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y, long long n_out, int H, int W, int Ho, int Wo, int ph, int pw) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(i % Wo);
        const int oh = (int)((i / Wo) % Ho);
        const long long n = i / ((long long)Wo * Ho);
        const float* p = x + n * H * W;
        float s = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int h = oh * 2 - ph + dy, w = ow * 2 - pw + dx;
                if (h >= 0 && h < H && w >= 0 && w < W) s += p[(long long)h * W + w];
            }
        y[i] = 0.25f * s;
    }
}

int avgpool2(const float* x, float* y, int N, int H, int W, cudaStream_t st) {
    const int ph = H % 2, pw = W % 2;
    const int Ho = (H + 2 * ph - 2) / 2 + 1, Wo = (W + 2 * pw - 2) / 2 + 1;
    const long long n = (long long)N * Ho * Wo;
    if (n <= 0) return VAE2_OK;
    avgpool2_kernel<<<stream_grid(n, 256), 256, 0, st>>>(x, y, n, H, W, Ho, Wo, ph, pw);
    return check_launch();
}

}  // namespace vae2
