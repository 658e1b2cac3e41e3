// Fused Adam over a flat fp32 parameter arena (torch.optim.Adam semantics as the reference
// configures it at tools/train.py:251-261: betas (0.9, 0.999), eps 1e-8, no amsgrad).
// One launch updates every parameter of a network; `step_dev` is the device-resident step
// counter (incremented by the caller's plan), `grad_scale` folds the data-parallel 1/world.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, const long long* __restrict__ step_dev,
            float grad_scale) {
    const float step = (float)(*step_dev);
    const float bc1 = 1.f - powf(b1, step);
    const float bc2s = sqrtf(1.f - powf(b2, step));
    const float step_size = lr / bc1;
    const long long n4 = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = fmaf(wd, pa[k], ga[k] * grad_scale);
            ma[k] = fmaf(b1, ma[k], (1.f - b1) * gr);
            va[k] = fmaf(b2, va[k], (1.f - b2) * gr * gr);
            pa[k] -= step_size * ma[k] / (sqrtf(va[k]) / bc2s + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gr = fmaf(wd, p[i], g[i] * grad_scale);
        m[i] = fmaf(b1, m[i], (1.f - b1) * gr);
        v[i] = fmaf(b2, v[i], (1.f - b2) * gr * gr);
        p[i] -= step_size * m[i] / (sqrtf(v[i]) / bc2s + eps);
    }
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float weight_decay, const long long* step_dev, float grad_scale, cudaStream_t st) {
    if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return VAE2_ERR_ARG;
    adam_kernel<<<stream_grid(n / 4 + 1, 256, 8), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale);
    return check_launch();
}

}  // namespace vae2
