// Internal C++ declarations of the kernel launchers (one translation unit per family).
// The public, reference-facing boundary is the C ABI in include/vae2_b200.h (api.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vae2 {

struct PackDesc {
    const float* w;        // OIHW fp32 parameter (pack source / unpack destination)
    float* wp;             // [tap][Cin_p][Cout_p] fp32 (pack dest / unpack source), may be null
    float* wpT;            // [tap][Cout_p][Cin_p] fp32, may be null
    __nv_bfloat16* wq;     // [tap][Cout_p][Cin_p] bf16, may be null
    __nv_bfloat16* wqT;    // [tap][Cin_p][Cout_p] bf16, may be null
    const int* cin_map;    // logical input channel -> physical lane, null = identity
    int Cout, Cin, k, Cin_p, Cout_p;
    int _pad;
};

struct ConvGeom {
    int B, H, W, Cin_p, ldx;     // input  [B][H][W][ldx], Cin_p lanes used
    int Ho, Wo, Cout_p, ldy;     // output [B][Ho][Wo][ldy], Cout_p lanes used
    int k, stride, pad;
};

// layout.cu
int nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int Cp, int H, int W, int ld,
                 int src_ctot, int src_coff, cudaStream_t st);
int nhwc_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W, int ld, int dst_ctot,
                 int dst_coff, int accumulate, cudaStream_t st);
int slice_copy(const void* src, void* dst, int dtype, long long P, int Cp, int ld_src, int ld_dst, int accumulate,
               cudaStream_t st);
int code_broadcast(const float* code, void* dst, int dtype, int B, int Z, int Zp, int H, int W, int ld, cudaStream_t st);
int spatial_sum(const void* x, void* out, int dtype, int out_fp32, int B, int HW, int C, int ld, int out_ld, float scale,
                int accumulate, cudaStream_t st);
int spatial_bcast(const void* g, void* dx, int dtype, int g_fp32, int B, int HW, int C, int Cp, int ld, int g_ld, float scale,
                  int accumulate, cudaStream_t st);
int pack_weights(const PackDesc* descs_dev, int n, cudaStream_t st);
int unpack_wgrad(const PackDesc* descs_dev, int n, int accumulate, cudaStream_t st);

// conv_simt.cu  (CUDA-core fp32-accumulate implicit GEMM; exact-fp32 path and cross-check)
int conv_fwd_simt(const void* x, const float* wp, const float* bias, void* y, int dtype, const ConvGeom& g, cudaStream_t st);
int conv_dgrad_simt(const void* dy, const float* wpT, void* dx, int dtype, const ConvGeom& g, int accumulate, cudaStream_t st);
int conv_wgrad_simt(const void* x, const void* dy, float* dwp, int dtype, const ConvGeom& g, cudaStream_t st);
int bias_grad(const void* dy, float* dbias, int dtype, long long P, int C, int ld, int accumulate, cudaStream_t st);

// bn.cu
// conv_direct.cu (fp32 register-tiled direct conv; VAE2_ERR_UNSUPPORTED -> caller falls back to the implicit GEMM)
int conv_fwd_direct(const float* x, const float* wp, const float* bias, float* y, const ConvGeom& g, cudaStream_t st);
int conv_wgrad_direct(const float* x, const float* dy, float* dwp, const ConvGeom& g, cudaStream_t st);
int conv_dgrad_direct(const float* dy, const float* wpT, float* dx, const ConvGeom& g, int accumulate, cudaStream_t st);

int bn_stats(const void* y, float* partials, int* n_partials_out, int dtype, long long P, int Cp, int ld, cudaStream_t st);
int bn_stats_max_partials();
int bn_merge(const float* partials, int n_partials, int Cp, float* merged, cudaStream_t st);
int bn_finalize(const float* parts, int n_parts, int C, int Cp, const float* gamma, const float* beta,
                float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                float* mean, float* invstd, float* scale, float* shift, cudaStream_t st, long long part_stride = 0);
int bn_eval_coeffs(int C, int Cp, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float eps, float* scale, float* shift, cudaStream_t st);
int bn_apply(const void* y, const void* res, void* out, int dtype, long long P, int Cp, int ld_y, int ld_res, int ld_out,
             const float* scale, const float* shift, int relu, cudaStream_t st);
int bn_bwd_reduce(const void* g, const void* a, const void* y, float* partials, int* n_partials_out, int dtype,
                  long long P, int Cp, int ld_g, int ld_a, int ld_y, const float* mean, const float* invstd, int relu,
                  cudaStream_t st);
void bn_debug_set_prof(unsigned long long* p);
int bn_fwd_fused(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                 int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale,
                 float* shift, int relu, cudaStream_t st);
int bn_bwd_fused(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                 long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                 const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                 int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, cudaStream_t st);
int bn_fwd_fused_groups(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                        int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd,
                        float* scale, float* shift, int relu, int groups, int stat_stride, cudaStream_t st);
int bn_bwd_fused_groups(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                        long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                        const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                        int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                        int stat_stride, cudaStream_t st);
int bn_sync_fwd_stats(const void* y, float* partials, int dtype, long long P, int C, int Cp, int ld_y, int groups, float* msg,
                      cudaStream_t st);
int bn_sync_fwd_apply(const void* y, const void* res, void* out, int dtype, long long P, int C, int Cp, int ld_y, int ld_res,
                      int ld_out, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      long long* nbt, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                      int relu, int groups, int stat_stride, const float* gathered, int n_parts, long long part_stride,
                      cudaStream_t st);
int bn_sync_bwd(int phase, const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups, int stat_stride,
                float* msg, const float* gsum, float inv_count, cudaStream_t st);
int bn_peer_setup(int world, int rank, void* const* bases, unsigned int* seq, int* err);
long long bn_peer_slot_words(int world, int groups, int Cp, int backward);
int bn_fwd_fused_peer(const void* y, const void* res, void* out, float* partials, int dtype, long long P, int C, int Cp,
                      int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* nbt, float momentum, float eps, float* mean, float* invstd,
                      float* scale, float* shift, int relu, int groups, int stat_stride, long long slot_word, int seq_index,
                      cudaStream_t st);
int bn_bwd_fused_peer(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                      long long P, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                      const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                      int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                      int stat_stride, float inv_count, long long slot_word, int seq_index, cudaStream_t st);
int ipc_alloc(long long bytes, void** ptr, void* handle64);
int ipc_open(const void* handle64, void** ptr);
int ipc_close(void* ptr);
int ipc_free(void* ptr);
int bn_bwd_finalize(const float* partials, int n_partials, int C, int Cp, float* sums, cudaStream_t st);
int bn_bwd_coeffs(const float* sums, int C, int Cp, float inv_count, float* dgamma, float* dbeta, int accumulate_param,
                  const float* sums_for_param, float* c1, float* c2, cudaStream_t st);
int bn_bwd_elemt(const void* g, const void* a, const void* y, void* dy, void* dres, int dtype, long long P, int Cp,
                 int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean, const float* invstd,
                 const float* scale, const float* c1, const float* c2, int relu, int acc_dy, int acc_dres,
                 cudaStream_t st);

// fuse.cu
struct FuseSrc { const void* ptr; int H, W, ld; };
int fuse_sum(const FuseSrc* srcs, int nsrc, void* out, int dtype, int B, int H, int W, int Cp, int ld_out, int relu,
             cudaStream_t st);
struct FuseDst { void* ptr; int ld; int accumulate; };
int fuse_bwd_same(const void* g, const void* out, const FuseDst* dsts, int ndst, int dtype, long long P, int Cp,
                  int ld_g, int ld_out, int relu, cudaStream_t st);
int fuse_bwd_up(const void* g, const void* out, void* gsrc, int dtype, int B, int H, int W, int Hs, int Ws, int Cp,
                int ld_g, int ld_out, int ld_gsrc, int relu, int accumulate, cudaStream_t st);

// elbo.cu
struct ElboSeg {
    int kind;            // 0 = L1, 1 = reparam+KL, 2 = LSGAN
    int slot;            // which accumulator the term adds into
    const float* a;      // L1: predict (NCHW fp32) | KL: eps (NCHW fp32) or null | GAN: sample (NCHW fp32)
    const float* b;      // L1: target           | KL: muvar (NCHW fp32 [B,2Z,H,W])
    float* out;          // KL: z out (NCHW fp32), may be null
    float target;        // GAN: 1 (real) / 0 (fake)
    float scale;         // term multiplier (1/B, 0.5/B ...)
    int Z, HW;           // KL: latent channels and pixels per map
    long long n;         // elements (L1/GAN: numel ; KL: B*Z*HW)
    int prior;           // KL: 1 = z = eps (prior sampling, utils.py:88-90)
    int _pad;
};
int elbo_terms(const ElboSeg* segs_dev, int nseg, float* acc, int nslots, int* nonfinite, cudaStream_t st);
int elbo_acc_floats();
struct ElboBwdSeg {
    int kind;            // 0 = L1, 1 = reparam+KL, 2 = LSGAN
    int _pad0;
    const float* a;      // as forward
    const float* b;
    const float* gz;     // KL: upstream dL/dz (NCHW fp32) or null
    float* grad;         // L1/GAN: d/d predict|sample ; KL: d/d muvar
    const float* gout;   // device scalar: upstream gradient of the term's loss value (or null = 1)
    float target, scale; // as forward; scale already includes lambda
    int Z, HW;
    long long n;
    int accumulate;
    int prior;
};
int elbo_terms_bwd(const ElboBwdSeg* segs_dev, int nseg, cudaStream_t st);

// metrics.cu (caller-side: device input pipeline, inference metrics)
int clip_u8_to_nchw(const uint8_t* src, float* dst, int B, int L, int H, int W, cudaStream_t st);
int to_image(const float* x, float* im, long long n, int HW, cudaStream_t st);
int frame_metrics(const float* pred, const float* gt, double* out, int R, int F, int Bg, int frame_elems, cudaStream_t st);
int ssim_level(const float* X, const float* Y, double* out, int N, int Ny, int H, int W, float data_range, cudaStream_t st);
int avgpool2(const float* x, float* y, int N, int H, int W, cudaStream_t st);

// adam.cu
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float weight_decay, const long long* step_dev, float grad_scale, cudaStream_t st);

// conv_tc.cu (tcgen05 / TMEM / TMA implicit GEMM, bf16)
int conv_fwd_tc(const void* x, const void* wq, const float* bias, void* y, const ConvGeom& g, float* stats_partials,
                cudaStream_t st);
int conv_tc_supported(const ConvGeom& g);
long long conv_wgrad_tc_workspace(const ConvGeom& g);
int conv_wgrad_tc(const void* x, const void* dy, float* dwp, float* ws, const ConvGeom& g, cudaStream_t st);
int conv_dgrad_tc(const void* dy, const void* wqT, void* dx, const ConvGeom& g, int accumulate, cudaStream_t st);
long long conv_wgrad_f32x2_workspace(const ConvGeom& g);      // bytes
int conv_wgrad_f32x2(const float* x, const float* dy, float* dwp, void* workspace, const ConvGeom& g, cudaStream_t st);

// conv_f32x3.cu (fp32-accurate tcgen05 convolution by an exact 3-way bf16 split: forward and data gradient of the fp32 path)
struct Tf32PackDesc {
    const float* w;        // OIHW fp32 parameter
    const int* cin_map;    // logical input channel -> physical lane, null = identity
    void* fwd;             // bf16 planes [3][tap][Nf][Kf], or null
    void* bwd;             // bf16 planes [3][tap][NfT][KfT], or null
    int Cout, Cin, k, Nf, Kf, NfT, KfT, _pad;
};
int conv_tf32_supported(const ConvGeom& g);
void conv_tf32_dims(const ConvGeom& g, int* Nf, int* Kf, int* NfT, int* KfT);
int pack_weights_tf32(const Tf32PackDesc* descs_dev, int n, cudaStream_t st);
int conv_fwd_tf32(const void* x, const void* w_fwd, const float* bias, void* y, const ConvGeom& g, cudaStream_t st);
int conv_dgrad_tf32(const void* dy, const void* w_bwd, void* dx, const ConvGeom& g, int accumulate, cudaStream_t st);

}  // namespace vae2
