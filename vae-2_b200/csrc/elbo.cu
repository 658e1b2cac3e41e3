// Fused ELBO terms: reparameterisation z = mu + exp(logvar/2)*eps, KL(q||N(0,I)), L1
// reconstruction and LSGAN terms as ONE table-driven, vectorised elementwise + reduction
// kernel (reference: lib/utils/utils.py:78-119 and lib/core/criterion.py:61-103, which run
// ~20 small ATen kernels plus 14 isnan/isinf host syncs per step, utils.py:63-65).
//
// A launch processes a table of segments; every segment streams its operands once with
// 16-byte loads, accumulates its term in fp32 registers, and the per-CTA partials are written
// to partials[cta][slot]; a second tiny kernel adds them in a fixed order in fp64
// (deterministic).  Non-finite values of z / predictions are counted into `nonfinite[seg]`
// instead of synchronising the host.
//
// All operands here are the reference's boundary tensors: NCHW fp32, contiguous.
#include "common.cuh"
#include "kernels.h"

namespace vae2 {

constexpr int ELBO_THREADS = 256;
constexpr int ELBO_MAX_SLOTS = 8;

__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= 3.402823466e38f; }

__device__ __forceinline__ float l1_term(float a, float b, int& bad) {
    if (!finite_f(a)) ++bad;
    return fabsf(a - b);
}

// One reparam/KL element.  e indexes [B][Z][HW]; muvar is [B][2Z][HW].
__device__ __forceinline__ float kl_elem(const ElboSeg& s, long long e, int& bad) {
    const long long zhw = (long long)s.Z * s.HW;
    const long long b = e / zhw, rem = e - b * zhw;
    const float mu = s.b[b * 2 * zhw + rem];
    const float lv = s.b[b * 2 * zhw + zhw + rem];
    if (s.out != nullptr) {
        const float ep = s.a != nullptr ? s.a[e] : 0.f;
        const float z = s.prior ? ep : fmaf(expf(0.5f * lv), ep, mu);
        if (!finite_f(z)) ++bad;
        s.out[e] = z;
    }
    // expm1f(lv) - lv instead of expf(lv) - lv - 1: no cancellation against 1 when |lv| is small
    return 0.5f * (mu * mu + (expm1f(lv) - lv));
}

__global__ void __launch_bounds__(ELBO_THREADS)
elbo_terms_kernel(const ElboSeg* __restrict__ segs, int nseg, float* __restrict__ partials, int nslots,
                  int* __restrict__ nonfinite) {
    __shared__ float red[32];
    __shared__ float slot_acc[ELBO_MAX_SLOTS];
    if (threadIdx.x < ELBO_MAX_SLOTS) slot_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (int si = 0; si < nseg; ++si) {
        const ElboSeg s = segs[si];
        float acc = 0.f;
        int bad = 0;
        if (s.kind == 0) {            // L1: sum |a - b|
            const bool vec = (s.n % 4 == 0) && ((((uintptr_t)s.a | (uintptr_t)s.b) & 15) == 0);
            if (vec) {
                const float4* a4 = reinterpret_cast<const float4*>(s.a);
                const float4* b4 = reinterpret_cast<const float4*>(s.b);
                for (long long i = tid; i < s.n / 4; i += nthreads) {
                    const float4 a = a4[i], b = b4[i];
                    acc += l1_term(a.x, b.x, bad) + l1_term(a.y, b.y, bad) + l1_term(a.z, b.z, bad) + l1_term(a.w, b.w, bad);
                }
            } else {
                for (long long i = tid; i < s.n; i += nthreads) acc += l1_term(s.a[i], s.b[i], bad);
            }
        } else if (s.kind == 1) {     // reparam + KL
            // rows r = (b, z) of HW contiguous elements: eps / z at r*HW, mu at (b*2Z + z)*HW, logvar Z planes further on
            const bool vec = (s.HW % 4 == 0) && (s.n / 4 < 0x7fffffffLL) &&
                             ((((uintptr_t)s.a | (uintptr_t)s.b | (uintptr_t)s.out) & 15) == 0);
            if (vec) {
                const unsigned hw4 = (unsigned)(s.HW / 4), n4 = (unsigned)(s.n / 4), Zu = (unsigned)s.Z;
                const float4* e4 = reinterpret_cast<const float4*>(s.a);
                const float4* m4 = reinterpret_cast<const float4*>(s.b);
                float4* o4 = reinterpret_cast<float4*>(s.out);
                for (unsigned i = (unsigned)tid; i < n4; i += (unsigned)nthreads) {
                    const unsigned r = i / hw4, c = i - r * hw4;
                    const unsigned b = r / Zu, z = r - b * Zu;
                    const size_t im = (size_t)(b * 2 * Zu + z) * hw4 + c;
                    const float4 mu = m4[im], lv = m4[im + (size_t)Zu * hw4];
                    if (o4 != nullptr) {
                        const float4 ep = e4 != nullptr ? e4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 zz;
                        zz.x = s.prior ? ep.x : fmaf(expf(0.5f * lv.x), ep.x, mu.x);
                        zz.y = s.prior ? ep.y : fmaf(expf(0.5f * lv.y), ep.y, mu.y);
                        zz.z = s.prior ? ep.z : fmaf(expf(0.5f * lv.z), ep.z, mu.z);
                        zz.w = s.prior ? ep.w : fmaf(expf(0.5f * lv.w), ep.w, mu.w);
                        bad += (!finite_f(zz.x)) + (!finite_f(zz.y)) + (!finite_f(zz.z)) + (!finite_f(zz.w));
                        o4[i] = zz;
                    }
                    acc += 0.5f * (mu.x * mu.x + (expm1f(lv.x) - lv.x)) + 0.5f * (mu.y * mu.y + (expm1f(lv.y) - lv.y)) +
                           0.5f * (mu.z * mu.z + (expm1f(lv.z) - lv.z)) + 0.5f * (mu.w * mu.w + (expm1f(lv.w) - lv.w));
                }
            } else {
                for (long long i = tid; i < s.n; i += nthreads) acc += kl_elem(s, i, bad);
            }
        } else {                      // LSGAN: sum (a - target)^2
            if ((s.n % 4 == 0) && (((uintptr_t)s.a & 15) == 0)) {
                const float4* a4 = reinterpret_cast<const float4*>(s.a);
                for (long long i = tid; i < s.n / 4; i += nthreads) {
                    const float4 a = a4[i];
                    const float d0 = a.x - s.target, d1 = a.y - s.target, d2 = a.z - s.target, d3 = a.w - s.target;
                    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                }
            } else {
                for (long long i = tid; i < s.n; i += nthreads) {
                    const float d = s.a[i] - s.target;
                    acc = fmaf(d, d, acc);
                }
            }
        }
        acc = block_sum(acc * s.scale, red);
        if (threadIdx.x == 0) slot_acc[s.slot] += acc;
        if (bad && nonfinite != nullptr) atomicAdd(nonfinite + si, bad);
    }
    __syncthreads();
    if (threadIdx.x < nslots) partials[(long long)blockIdx.x * nslots + threadIdx.x] = slot_acc[threadIdx.x];
}

__global__ void elbo_reduce_kernel(const float* __restrict__ partials, int nblocks, int nslots, float* __restrict__ out) {
    const int s = threadIdx.x;
    if (s >= nslots) return;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += (double)partials[(long long)b * nslots + s];
    out[s] = (float)t;
}

// acc layout: [0, nslots) final sums, followed by grid*nslots floats of per-CTA partials.
static int elbo_grid() { return kNumSMs * 4; }

int elbo_terms(const ElboSeg* segs_dev, int nseg, float* acc, int nslots, int* nonfinite, cudaStream_t st) {
    if (nslots < 1 || nslots > ELBO_MAX_SLOTS || nseg < 1) return VAE2_ERR_ARG;
    const int grid = elbo_grid();
    float* partials = acc + ELBO_MAX_SLOTS;
    elbo_terms_kernel<<<grid, ELBO_THREADS, 0, st>>>(segs_dev, nseg, partials, nslots, nonfinite);
    if (int e = check_launch()) return e;
    elbo_reduce_kernel<<<1, 32, 0, st>>>(partials, grid, nslots, acc);
    return check_launch();
}

int elbo_acc_floats() { return ELBO_MAX_SLOTS + elbo_grid() * ELBO_MAX_SLOTS; }

// ---------------------------------------------------------------------------
// backward: elementwise gradients of every term (no reduction)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(ELBO_THREADS)
elbo_terms_bwd_kernel(const ElboBwdSeg* __restrict__ segs, int nseg) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (int si = 0; si < nseg; ++si) {
        const ElboBwdSeg s = segs[si];
        const float go = (s.gout != nullptr ? *s.gout : 1.f) * s.scale;
        if (s.kind == 0) {
            if ((s.n % 4 == 0) && ((((uintptr_t)s.a | (uintptr_t)s.b | (uintptr_t)s.grad) & 15) == 0)) {
                const float4* a4 = reinterpret_cast<const float4*>(s.a);
                const float4* b4 = reinterpret_cast<const float4*>(s.b);
                float4* g4 = reinterpret_cast<float4*>(s.grad);
                for (long long i = tid; i < s.n / 4; i += nthreads) {
                    const float4 a = a4[i], b = b4[i];
                    float4 gr;
                    gr.x = a.x > b.x ? go : (a.x < b.x ? -go : 0.f);
                    gr.y = a.y > b.y ? go : (a.y < b.y ? -go : 0.f);
                    gr.z = a.z > b.z ? go : (a.z < b.z ? -go : 0.f);
                    gr.w = a.w > b.w ? go : (a.w < b.w ? -go : 0.f);
                    if (s.accumulate) { const float4 o = g4[i]; gr.x += o.x; gr.y += o.y; gr.z += o.z; gr.w += o.w; }
                    g4[i] = gr;
                }
                continue;
            }
            for (long long i = tid; i < s.n; i += nthreads) {
                const float d = s.a[i] - s.b[i];
                const float gr = d > 0.f ? go : (d < 0.f ? -go : 0.f);
                s.grad[i] = s.accumulate ? s.grad[i] + gr : gr;
            }
        } else if (s.kind == 1) {
            const long long zhw = (long long)s.Z * s.HW;
            const bool vec = (s.HW % 4 == 0) && (s.n / 4 < 0x7fffffffLL) &&
                             ((((uintptr_t)s.a | (uintptr_t)s.b | (uintptr_t)s.gz | (uintptr_t)s.grad) & 15) == 0);
            if (vec) {
                const unsigned hw4 = (unsigned)(s.HW / 4), n4 = (unsigned)(s.n / 4), Zu = (unsigned)s.Z;
                const float4* e4 = reinterpret_cast<const float4*>(s.a);
                const float4* m4 = reinterpret_cast<const float4*>(s.b);
                const float4* z4 = reinterpret_cast<const float4*>(s.gz);
                float4* g4 = reinterpret_cast<float4*>(s.grad);
                const bool use_gz = s.gz != nullptr && !s.prior;
                for (unsigned i = (unsigned)tid; i < n4; i += (unsigned)nthreads) {
                    const unsigned r = i / hw4, c = i - r * hw4;
                    const unsigned b = r / Zu, z = r - b * Zu;
                    const size_t im = (size_t)(b * 2 * Zu + z) * hw4 + c, iv = im + (size_t)Zu * hw4;
                    const float4 mu = m4[im], lv = m4[iv];
                    float4 dmu = make_float4(go * mu.x, go * mu.y, go * mu.z, go * mu.w);
                    float4 dlv = make_float4(go * 0.5f * expm1f(lv.x), go * 0.5f * expm1f(lv.y), go * 0.5f * expm1f(lv.z),
                                             go * 0.5f * expm1f(lv.w));
                    if (use_gz) {
                        const float4 gz = z4[i], ep = e4[i];
                        dmu.x += gz.x; dmu.y += gz.y; dmu.z += gz.z; dmu.w += gz.w;
                        dlv.x = fmaf(gz.x * 0.5f * expf(0.5f * lv.x), ep.x, dlv.x);
                        dlv.y = fmaf(gz.y * 0.5f * expf(0.5f * lv.y), ep.y, dlv.y);
                        dlv.z = fmaf(gz.z * 0.5f * expf(0.5f * lv.z), ep.z, dlv.z);
                        dlv.w = fmaf(gz.w * 0.5f * expf(0.5f * lv.w), ep.w, dlv.w);
                    }
                    if (s.accumulate) {
                        const float4 a = g4[im], v = g4[iv];
                        dmu.x += a.x; dmu.y += a.y; dmu.z += a.z; dmu.w += a.w;
                        dlv.x += v.x; dlv.y += v.y; dlv.z += v.z; dlv.w += v.w;
                    }
                    g4[im] = dmu;
                    g4[iv] = dlv;
                }
                continue;
            }
            for (long long e = tid; e < s.n; e += nthreads) {
                const long long b = e / zhw, rem = e - b * zhw;
                const long long im = b * 2 * zhw + rem, iv = im + zhw;
                const float mu = s.b[im], lv = s.b[iv];
                float dmu = go * mu;
                float dlv = go * 0.5f * expm1f(lv);
                if (s.gz != nullptr && !s.prior) {
                    const float gz = s.gz[e];
                    dmu += gz;
                    dlv = fmaf(gz * 0.5f * expf(0.5f * lv), s.a[e], dlv);
                }
                if (s.accumulate) { dmu += s.grad[im]; dlv += s.grad[iv]; }
                s.grad[im] = dmu;
                s.grad[iv] = dlv;
            }
        } else {
            for (long long i = tid; i < s.n; i += nthreads) {
                const float gr = go * 2.f * (s.a[i] - s.target);
                s.grad[i] = s.accumulate ? s.grad[i] + gr : gr;
            }
        }
    }
}

int elbo_terms_bwd(const ElboBwdSeg* segs_dev, int nseg, cudaStream_t st) {
    if (nseg < 1) return VAE2_ERR_ARG;
    elbo_terms_bwd_kernel<<<kNumSMs * 8, ELBO_THREADS, 0, st>>>(segs_dev, nseg);
    return check_launch();
}

}  // namespace vae2
