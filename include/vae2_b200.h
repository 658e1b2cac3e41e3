/*
 * vae2_b200.h -- C ABI of the B200-native VAE^2 hot path (libvae2_b200.so).
 *
 * The reference has no FFI registry; its boundary for this path is the nn.Module surface
 * (lib/models/enc_hrnet.py:1185-1210, lib/utils/utils.py:39-155, lib/core/criterion.py:61-103)
 * and, for native ops, the free-function op convention of its vendored extension
 * (lib/models/sync_bn/inplace_abn/src/inplace_abn.cpp:7-75: one function per op, device
 * pointers borrowed from the caller, errors reported to Python as RuntimeError).  This header
 * is that convention as a plain C ABI: raw DEVICE pointers + sizes + a cudaStream_t, an int
 * status back (0 = ok).  No torch types.  Every entry point names the reference call site(s)
 * whose arithmetic it replaces.  The Python mirror of the reference interface
 * (vae-2_b200/lib/...) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Tensor convention: "act" tensors are channels-last [B][H][W][ld] in fp32 (dtype 0) or bf16
 * (dtype 1); Cp = padded channel lanes actually used (multiple of 4 / 8), ld >= Cp the pixel
 * pitch in elements (so a tensor may be a channel slice of a concat buffer).  Pad lanes are 0.
 * "nchw" tensors are the reference's own layout: contiguous fp32 [B][C][H][W].
 * All functions are asynchronous on `stream` and re-entrant (no global state).
 */
#ifndef VAE2_B200_H_
#define VAE2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAE2_ABI_VERSION 1

typedef void* vae2_stream_t; /* cudaStream_t */

/* status codes */
enum { VAE2_STATUS_OK = 0, VAE2_STATUS_BAD_ARG = 1, VAE2_STATUS_CUDA = 2, VAE2_STATUS_UNSUPPORTED = 3 };

int vae2_abi_version(void);
const char* vae2_status_string(int status);
/* last CUDA error string seen by this library on the calling thread's device */
const char* vae2_last_cuda_error(void);
/* name of the kernel the last convolution launcher on the calling thread dispatched to (profiling introspection) */
const char* vae2_last_kernel(void);

/* ---- layout: the reference's NCHW fp32 tensors <-> internal channels-last ------------------ */
/* xs[i].to(device) inputs, lib/core/function.py:487-489; torch.cat inputs, lib/utils/utils.py:77,105 */
int vae2_nchw_to_act(const float* src, void* dst, int dtype, int B, int C, int Cp, int H, int W, int ld,
                     int src_ctot, int src_coff, vae2_stream_t stream);
/* network outputs (predictions, mu/logvar maps), torch.cat of head outputs enc_hrnet.py:845 */
int vae2_act_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W, int ld, int dst_ctot,
                     int dst_coff, int accumulate, vae2_stream_t stream);
/* torch.cat along channels (enc_hrnet.py:825-826, 839, 885, 943) done as slice writes */
int vae2_slice_copy(const void* src, void* dst, int dtype, int64_t npix, int Cp, int ld_src, int ld_dst,
                    int accumulate, vae2_stream_t stream);
/* HighResolutionNet._gen_code_map, enc_hrnet.py:454-462 */
int vae2_code_broadcast(const float* code, void* dst, int dtype, int B, int Z, int Zp, int H, int W, int ld,
                        vae2_stream_t stream);

/* nn.AdaptiveAvgPool2d((1,1)) of the non-HD_Z posterior head (enc_hrnet.py:1023-1041) and the gradient of the spatial
 * repeat of a per-sample z (enc_hrnet.py:454-462):  out[b][c] (=|+=) scale * sum_p x[b][p][c]; out is fp32 (out_fp32=1)
 * or an act of x's dtype, row pitch out_ld elements */
int vae2_spatial_sum(const void* x, void* out, int dtype, int out_fp32, int B, int HW, int C, int ld, int out_ld, float scale,
                     int accumulate, vae2_stream_t stream);
/* its transpose: dx[b][p][c] (=|+=) scale * g[b][c] for c < C, 0 for the pad lanes C..Cp */
int vae2_spatial_bcast(const void* g, void* dx, int dtype, int g_fp32, int B, int HW, int C, int Cp, int ld, int g_ld,
                       float scale, int accumulate, vae2_stream_t stream);

/* ---- weights: OIHW nn.Conv2d parameters <-> GEMM operand layouts ---------------------------- */
typedef struct {
    const float* w;      /* OIHW fp32 (pack: source, unpack: destination gradient) */
    float* wp;           /* [tap][Cin_p][Cout_p] fp32 (pack: dest, unpack: source) or NULL */
    float* wpT;          /* [tap][Cout_p][Cin_p] fp32 or NULL */
    void* wq;            /* [tap][Cout_p][Cin_p] bf16 or NULL */
    void* wqT;           /* [tap][Cin_p][Cout_p] bf16 or NULL */
    const int32_t* cin_map; /* logical input channel -> physical lane (concat inputs), NULL = identity */
    int32_t Cout, Cin, k, Cin_p, Cout_p, reserved;
} vae2_pack_desc;        /* array lives in DEVICE memory */
int vae2_pack_weights(const vae2_pack_desc* descs_dev, int n, vae2_stream_t stream);
int vae2_unpack_wgrad(const vae2_pack_desc* descs_dev, int n, int accumulate, vae2_stream_t stream);

/* ---- convolution: every nn.Conv2d site of enc_hrnet.py (3x3 s1/s2 p1, 1x1) ------------------ */
typedef struct {
    int32_t B, H, W, Cin_p, ldx;
    int32_t Ho, Wo, Cout_p, ldy;
    int32_t k, stride, pad;
} vae2_conv_geom;
/* engine: 0 = CUDA-core fp32 FMA (exact fp32), 1 = tcgen05/TMEM/TMA (bf16 operands, dtype must be 1),
 *         2 = tcgen05 with an exact 3-way bf16 split of every fp32 operand and 6 accumulated products (fp32
 *             tensors, fp32-level accuracy; forward and dgrad only; w_packed from vae2_pack_weights_tf32) */
int vae2_conv2d_fwd(const void* x, const void* w_packed, const float* bias, void* y, int dtype,
                    const vae2_conv_geom* g, int engine, vae2_stream_t stream);
int vae2_conv2d_dgrad(const void* dy, const void* w_packed_t, void* dx, int dtype, const vae2_conv_geom* g,
                      int accumulate, int engine, vae2_stream_t stream);
/* dw_packed fp32 [tap][Cin_p][Cout_p], must be zeroed by the caller (split-K accumulation) */
int vae2_conv2d_wgrad(const void* x, const void* dy, float* dw_packed, int dtype, const vae2_conv_geom* g,
                      int engine, vae2_stream_t stream);
int vae2_bias_grad(const void* dy, float* dbias, int dtype, int64_t npix, int C, int ld, int accumulate,
                   vae2_stream_t stream);
int vae2_conv2d_tc_supported(const vae2_conv_geom* g);
/* engine-2 weights: three bf16 planes, [3][tap][Nf][Kf] (forward) and [3][tap][NfT][KfT] (data gradient),
 * zero-initialised by the caller; dims from vae2_conv2d_tf32_dims */
typedef struct {
    const float* w;
    const int32_t* cin_map;
    void* fwd;
    void* bwd;
    int32_t Cout, Cin, k, Nf, Kf, NfT, KfT, reserved;
} vae2_tf32_pack_desc;   /* DEVICE array */
int vae2_conv2d_tf32_supported(const vae2_conv_geom* g);
void vae2_conv2d_tf32_dims(const vae2_conv_geom* g, int* Nf, int* Kf, int* NfT, int* KfT);
int vae2_pack_weights_tf32(const vae2_tf32_pack_desc* descs_dev, int n, vae2_stream_t stream);
/* tensor-core weight gradient (bf16 act, stride 1): dw_packed is OVERWRITTEN; `workspace` holds the split-K
 * partials, at least vae2_conv2d_wgrad_tc_workspace(g) floats (negative = shape unsupported, use engine 0) */
long long vae2_conv2d_wgrad_tc_workspace(const vae2_conv_geom* g);
int vae2_conv2d_wgrad_tc(const void* x, const void* dy, float* dw_packed, float* workspace, const vae2_conv_geom* g,
                         vae2_stream_t stream);

/* fp32 tensor-core weight gradient (engine-2 companion): x and dy (fp32 act) are split into two bf16 planes each and the
 * three leading plane products run on the tcgen05 weight-gradient kernel with separate accumulators; dw_packed fp32
 * [tap][Cin_p][Cout_p] is OVERWRITTEN.  `workspace`: vae2_conv2d_wgrad_f32x2_workspace(g) BYTES (negative = unsupported) */
long long vae2_conv2d_wgrad_f32x2_workspace(const vae2_conv_geom* g);
int vae2_conv2d_wgrad_f32x2(const float* x, const float* dy, float* dw_packed, void* workspace, const vae2_conv_geom* g,
                            vae2_stream_t stream);

/* ---- batch norm: BatchNorm2d(momentum=0.01) / SyncBatchNorm, enc_hrnet.py:22-23, train.py:217 - */
int vae2_bn_max_partials(void);
/* per-CTA Welford partials [n_partials][3][Cp] = (count, mean, M2); *n_partials is a HOST out */
int vae2_bn_stats(const void* y, float* partials, int* n_partials, int dtype, int64_t npix, int Cp, int ld,
                  vae2_stream_t stream);
/* merge partial sets into one [3][Cp] set (the per-rank message of the SyncBN all-gather) */
int vae2_bn_merge(const float* partials, int n_partials, int Cp, float* merged, vae2_stream_t stream);
/* merge + mean/invstd + scale=gamma*invstd, shift=beta-mean*scale + running stats + num_batches_tracked */
int vae2_bn_finalize(const float* partials, int n_partials, int C, int Cp, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                     float eps, float* mean, float* invstd, float* scale, float* shift, vae2_stream_t stream);
/* same, with the partial sets `part_stride` floats apart: several BNs of one SyncBN group share ONE all-gather
 * message, so rank r's partial of this BN sits at partials + r*part_stride */
int vae2_bn_finalize_strided(const float* partials, int n_partials, int64_t part_stride, int C, int Cp,
                             const float* gamma, const float* beta, float* running_mean, float* running_var,
                             int64_t* num_batches_tracked, float momentum, float eps, float* mean, float* invstd,
                             float* scale, float* shift, vae2_stream_t stream);
int vae2_bn_eval_coeffs(int C, int Cp, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, vae2_stream_t stream);
/* out = [relu](scale*y + shift [+ res])   (BasicBlock/Bottleneck tails, enc_hrnet.py:46-62, 83-103) */
int vae2_bn_apply(const void* y, const void* res, void* out, int dtype, int64_t npix, int Cp, int ld_y,
                  int ld_res, int ld_out, const float* scale, const float* shift, int relu, vae2_stream_t stream);
int vae2_bn_bwd_reduce(const void* g, const void* a, const void* y, float* partials, int* n_partials, int dtype,
                       int64_t npix, int Cp, int ld_g, int ld_a, int ld_y, const float* mean, const float* invstd,
                       int relu, vae2_stream_t stream);
int vae2_bn_bwd_finalize(const float* partials, int n_partials, int C, int Cp, float* sums, vae2_stream_t stream);
int vae2_bn_bwd_coeffs(const float* sums_global, int C, int Cp, float inv_count, float* dgamma, float* dbeta,
                       int accumulate_param, const float* sums_local, float* c1, float* c2, vae2_stream_t stream);
int vae2_bn_bwd_elemt(const void* g, const void* a, const void* y, void* dy, void* dres, int dtype, int64_t npix,
                      int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                      const float* invstd, const float* scale, const float* c1, const float* c2, int relu,
                      int acc_dy, int acc_dres, vae2_stream_t stream);
/* Single-rank training BN in ONE cooperative launch per direction (statistics | grid barrier | finalize |
 * grid barrier | elementwise).  Same results as stats+finalize+apply resp. bwd_reduce+bwd_finalize+bwd_coeffs+
 * bwd_elemt above; `partials` is scratch of vae2_bn_max_partials()*3*Cp floats.  SyncBN keeps the split entry
 * points because its collective sits between the phases.  bwd `relu`: 0 none, 1 mask read from the stored
 * activation `a`, 2 mask recomputed as fma(y, scale, shift) > 0 (BN+ReLU without a residual: `a` is not read). */
/* debug aid: when set (device buffer of 6 x uint64), CTA 0 of the fused kernels records %globaltimer at its phase
 * boundaries (start | stats done | barrier 1 passed | finalize done | barrier 2 passed | end); null disables */
int vae2_debug_bn_phase_times(uint64_t* device_buf6);
int vae2_bn_fwd_fused(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C,
                      int Cp, int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                      float eps, float* mean, float* invstd, float* scale, float* shift, int relu,
                      vae2_stream_t stream);
int vae2_bn_bwd_fused(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                      int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                      const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                      float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                      vae2_stream_t stream);

/* The same with `groups` statistics groups stacked along the batch axis (the reference calls a discriminator once per
 * frame / per real-fake input, lib/utils/utils.py:114-119, 259-267; here those calls are one stacked pass): group g covers
 * npix pixels, tensors of consecutive groups follow each other in memory, statistic outputs (mean, invstd, scale, shift,
 * c1, c2) of group g sit g*stat_stride floats after group 0's.  Each group is normalised with its own batch statistics;
 * running statistics receive the `groups` momentum updates in group order, num_batches_tracked advances by `groups`,
 * d(gamma) / d(beta) are summed over the groups -- exactly what `groups` sequential module calls produce. */
int vae2_bn_fwd_fused_groups(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C,
                             int Cp, int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                             float eps, float* mean, float* invstd, float* scale, float* shift, int relu, int groups,
                             int stat_stride, vae2_stream_t stream);
int vae2_bn_bwd_fused_groups(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                             int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                             const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                             float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                             int groups, int stat_stride, vae2_stream_t stream);

/* SyncBatchNorm (tools/train.py:217; torch nn/modules/_functions.py:39-122) as TWO cooperative launches per direction around
 * the NCCL collective the host issues between them, every statistics group of a stacked pass in the same launch:
 *   forward : vae2_bn_sync_fwd_stats  -> msg[g][3][Cp] = rank-local (count, mean, M2)      | all-gather of the messages |
 *             vae2_bn_sync_fwd_apply  <- gathered: set r of group g at gathered + r*part_stride + g*3*Cp (n_parts = world)
 *   backward: vae2_bn_sync_bwd(1,..)  -> msg[g][2][Cp] = rank-local (sum dyb, sum dyb*xhat), d(gamma)/d(beta) from them
 *             | all-reduce | vae2_bn_sync_bwd(2,..) <- gsum[g][2][Cp], inv_count = 1 / (npix * world) */
int vae2_bn_sync_fwd_stats(const void* y, float* partials, int dtype, int64_t npix, int C, int Cp, int ld_y, int groups,
                           float* msg, vae2_stream_t stream);
int vae2_bn_sync_fwd_apply(const void* y, const void* res, void* out, int dtype, int64_t npix, int C, int Cp, int ld_y,
                           int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* mean,
                           float* invstd, float* scale, float* shift, int relu, int groups, int stat_stride,
                           const float* gathered, int n_parts, int64_t part_stride, vae2_stream_t stream);
int vae2_bn_sync_bwd(int phase, const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                     int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                     const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                     int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                     int stat_stride, float* msg, const float* gsum, float inv_count, vae2_stream_t stream);

/* SyncBatchNorm with the cross-rank exchange INSIDE the launch, over NVLink peer memory (no NCCL call per layer): the same
 * single cooperative launch per direction as on one GPU; the CTA that finalizes a channel stores its rank-local
 * (count, mean, M2) / (sum dyb, sum dyb*xhat) into every rank's mailbox with P2P stores and merges what the other ranks
 * stored into its own, in rank order (bit-identical statistics on every rank).  Replaces torch's SyncBatchNorm collectives
 * (nn/modules/_functions.py:39-122: all_gather of the statistics, all_reduce of the backward sums) for tools/train.py:217.
 *   vae2_ipc_alloc / vae2_ipc_open : one zeroed device allocation per process + its 64-byte CUDA IPC handle; peers map it.
 *   vae2_bn_peer_setup             : bases[r] = rank r's mailbox as mapped in this process, seq = zeroed uint32 per op,
 *                                    err = zeroed int (set when a peer did not answer within ~60 s; the host must raise).
 *   vae2_bn_peer_slot_words        : 8-byte words an op needs in the mailbox; the host gives every op (and direction) its own
 *                                    slot_word offset and seq_index, identical on every rank.
 * Gradient all-reduce kernels must not be in flight while these launches wait for their peers. */
int vae2_ipc_alloc(int64_t bytes, void** ptr, void* handle64);
int vae2_ipc_open(const void* handle64, void** ptr);
int vae2_ipc_close(void* ptr);
int vae2_ipc_free(void* ptr);
int vae2_bn_peer_setup(int world, int rank, void* const* bases, uint32_t* seq, int* err);
int64_t vae2_bn_peer_slot_words(int world, int groups, int Cp, int backward);
int vae2_bn_fwd_fused_peer(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C,
                           int Cp, int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                           float eps, float* mean, float* invstd, float* scale, float* shift, int relu, int groups,
                           int stat_stride, int64_t slot_word, int seq_index, vae2_stream_t stream);
int vae2_bn_bwd_fused_peer(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                           int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                           const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                           float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                           int groups, int stat_stride, float inv_count, int64_t slot_word, int seq_index,
                           vae2_stream_t stream);

/* ---- branch fusion / upsampling: HighResolutionModule.forward enc_hrnet.py:233-248, :833-839 -- */
typedef struct { const void* ptr; int32_t H, W, ld; } vae2_fuse_src;        /* HOST array */
typedef struct { void* ptr; int32_t ld, accumulate; } vae2_fuse_dst;         /* HOST array */
int vae2_fuse_sum(const vae2_fuse_src* srcs, int nsrc, void* out, int dtype, int B, int H, int W, int Cp,
                  int ld_out, int relu, vae2_stream_t stream);
int vae2_fuse_bwd_same(const void* g, const void* out, const vae2_fuse_dst* dsts, int ndst, int dtype,
                       int64_t npix, int Cp, int ld_g, int ld_out, int relu, vae2_stream_t stream);
int vae2_fuse_bwd_up(const void* g, const void* out, void* gsrc, int dtype, int B, int H, int W, int Hs, int Ws,
                     int Cp, int ld_g, int ld_out, int ld_gsrc, int relu, int accumulate, vae2_stream_t stream);

/* ---- ELBO terms: utils.py:78-119 (reparam, finite check), criterion.py:61-103 (L1, KL, LSGAN) - */
typedef struct {
    int32_t kind, slot;   /* kind 0 = L1, 1 = reparam+KL, 2 = LSGAN; slot = output accumulator */
    const float* a;       /* L1: predict | KL: eps (or NULL) | GAN: sample        (nchw fp32) */
    const float* b;       /* L1: target  | KL: muvar [B,2Z,H,W]                   (nchw fp32) */
    float* out;           /* KL: z [B,Z,H,W] or NULL */
    float target, scale;
    int32_t Z, HW;
    int64_t n;
    int32_t prior, reserved;
} vae2_elbo_seg;          /* DEVICE array */
typedef struct {
    int32_t kind, reserved0;
    const float* a;
    const float* b;
    const float* gz;      /* KL: dL/dz or NULL */
    float* grad;          /* L1/GAN: d/d a ; KL: d/d muvar */
    const float* gout;    /* device scalar upstream gradient of the term (NULL = 1) */
    float target, scale;
    int32_t Z, HW;
    int64_t n;
    int32_t accumulate, prior;
} vae2_elbo_bwd_seg;      /* DEVICE array */
int vae2_elbo_acc_floats(void);  /* floats the caller must provide in `acc` */
/* acc[0..nslots) receive the term sums; nonfinite[seg] counts inf/nan in z / predictions */
int vae2_elbo_terms(const vae2_elbo_seg* segs_dev, int nseg, float* acc, int nslots, int32_t* nonfinite,
                    vae2_stream_t stream);
int vae2_elbo_terms_bwd(const vae2_elbo_bwd_seg* segs_dev, int nseg, vae2_stream_t stream);

/* ---- callers either side of the path ------------------------------------ */
/* dataset frame preparation on the device (lib/datasets/cityscapes.py:300-326): uint8 RGB frames [B][L][H][W][3] ->
 * fp32 nchw [B][3L][H][W], (v/255 - mean)/std with the ImageNet constants the reference uses */
int vae2_clip_u8_to_nchw(const uint8_t* frames, float* dst, int B, int L, int H, int W, vae2_stream_t stream);
/* `_to_image(x, is_uint8=False)` of the inference driver (lib/core/function.py:87-98): (x*std+mean)*255 clipped to
 * [0,255]; x nchw with RGB triplets along channels, n = total elements, HW = pixels per plane */
int vae2_to_image(const float* x, float* im, int64_t n, int HW, vae2_stream_t stream);
/* per (row r, frame f): out[(r*F+f)*2] = sum |pred - gt|, [..+1] = sum (pred - gt)^2 over the frame's 3*H*W values
 * (function.py:262-263 recon loss and PSNR inputs); image planes [R][F][frame_elems], gt row = r % Bg */
int vae2_frame_metrics(const float* pred_im, const float* gt_im, double* out, int R, int F, int Bg, int frame_elems,
                       vae2_stream_t stream);
/* one SSIM level as pytorch_msssim computes it (function.py:244-261; 11-tap Gaussian sigma 1.5, valid, K=(0.01,0.03)):
 * planes X [N][H][W] vs Y [Ny][H][W] (plane n against n % Ny); out[n*2] = sum ssim_map, out[n*2+1] = sum cs_map */
int vae2_ssim_level(const float* X, const float* Y, double* out, int N, int Ny, int H, int W, float data_range,
                    vae2_stream_t stream);
/* F.avg_pool2d(kernel 2, padding (H%2, W%2)) between MS-SSIM levels; planes [N][H][W] -> [N][Ho][Wo] */
int vae2_avgpool2(const float* x, float* y, int N, int H, int W, vae2_stream_t stream);

/* ---- optimizer: torch.optim.Adam as configured at tools/train.py:251-261 --------------------- */
int vae2_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, const int64_t* step_dev, float grad_scale, vae2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAE2_B200_H_ */
