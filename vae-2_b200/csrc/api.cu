// extern "C" boundary: include/vae2_b200.h  ->  namespace vae2 launchers.
#include "../../include/vae2_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace vae2;

namespace vae2 { thread_local const char* g_last_kernel = ""; }

static_assert(sizeof(vae2_pack_desc) == sizeof(PackDesc), "pack desc ABI");
static_assert(sizeof(vae2_conv_geom) == sizeof(ConvGeom), "conv geom ABI");
static_assert(sizeof(vae2_fuse_src) == sizeof(FuseSrc), "fuse src ABI");
static_assert(sizeof(vae2_fuse_dst) == sizeof(FuseDst), "fuse dst ABI");
static_assert(sizeof(vae2_elbo_seg) == sizeof(ElboSeg), "elbo seg ABI");
static_assert(sizeof(vae2_elbo_bwd_seg) == sizeof(ElboBwdSeg), "elbo bwd seg ABI");
static_assert(sizeof(vae2_tf32_pack_desc) == sizeof(Tf32PackDesc), "tf32 pack desc ABI");

static inline cudaStream_t S(vae2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline const ConvGeom& G(const vae2_conv_geom* g) { return *reinterpret_cast<const ConvGeom*>(g); }

extern "C" {

int vae2_abi_version(void) { return VAE2_ABI_VERSION; }

const char* vae2_status_string(int status) {
    switch (status) {
        case VAE2_STATUS_OK: return "ok";
        case VAE2_STATUS_BAD_ARG: return "bad argument (alignment / size contract violated)";
        case VAE2_STATUS_CUDA: return "CUDA launch error";
        case VAE2_STATUS_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown status";
}

const char* vae2_last_cuda_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }
const char* vae2_last_kernel(void) { return g_last_kernel; }

int vae2_nchw_to_act(const float* src, void* dst, int dtype, int B, int C, int Cp, int H, int W, int ld, int src_ctot,
                     int src_coff, vae2_stream_t stream) {
    return nchw_to_nhwc(src, dst, dtype, B, C, Cp, H, W, ld, src_ctot, src_coff, S(stream));
}
int vae2_act_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W, int ld, int dst_ctot,
                     int dst_coff, int accumulate, vae2_stream_t stream) {
    return nhwc_to_nchw(src, dst, dtype, B, C, H, W, ld, dst_ctot, dst_coff, accumulate, S(stream));
}
int vae2_slice_copy(const void* src, void* dst, int dtype, int64_t npix, int Cp, int ld_src, int ld_dst, int accumulate,
                    vae2_stream_t stream) {
    return slice_copy(src, dst, dtype, npix, Cp, ld_src, ld_dst, accumulate, S(stream));
}
int vae2_code_broadcast(const float* code, void* dst, int dtype, int B, int Z, int Zp, int H, int W, int ld,
                        vae2_stream_t stream) {
    return code_broadcast(code, dst, dtype, B, Z, Zp, H, W, ld, S(stream));
}
int vae2_spatial_sum(const void* x, void* out, int dtype, int out_fp32, int B, int HW, int C, int ld, int out_ld, float scale,
                     int accumulate, vae2_stream_t stream) {
    return spatial_sum(x, out, dtype, out_fp32, B, HW, C, ld, out_ld, scale, accumulate, S(stream));
}
int vae2_spatial_bcast(const void* g, void* dx, int dtype, int g_fp32, int B, int HW, int C, int Cp, int ld, int g_ld,
                       float scale, int accumulate, vae2_stream_t stream) {
    return spatial_bcast(g, dx, dtype, g_fp32, B, HW, C, Cp, ld, g_ld, scale, accumulate, S(stream));
}
int vae2_pack_weights(const vae2_pack_desc* d, int n, vae2_stream_t stream) {
    return pack_weights(reinterpret_cast<const PackDesc*>(d), n, S(stream));
}
int vae2_unpack_wgrad(const vae2_pack_desc* d, int n, int accumulate, vae2_stream_t stream) {
    return unpack_wgrad(reinterpret_cast<const PackDesc*>(d), n, accumulate, S(stream));
}

int vae2_conv2d_fwd(const void* x, const void* w_packed, const float* bias, void* y, int dtype, const vae2_conv_geom* g,
                    int engine, vae2_stream_t stream) {
    if (engine == 1) {
        if (dtype != VAE2_DT_BF16) return VAE2_ERR_ARG;
        return conv_fwd_tc(x, w_packed, bias, y, G(g), nullptr, S(stream));
    }
    if (engine == 2) {
        if (dtype != VAE2_DT_F32) return VAE2_ERR_ARG;
        return conv_fwd_tf32(x, w_packed, bias, y, G(g), S(stream));
    }
    return conv_fwd_simt(x, reinterpret_cast<const float*>(w_packed), bias, y, dtype, G(g), S(stream));
}
int vae2_conv2d_dgrad(const void* dy, const void* w_packed_t, void* dx, int dtype, const vae2_conv_geom* g, int accumulate,
                      int engine, vae2_stream_t stream) {
    if (engine == 1) {
        if (dtype != VAE2_DT_BF16) return VAE2_ERR_ARG;
        return conv_dgrad_tc(dy, w_packed_t, dx, G(g), accumulate, S(stream));
    }
    if (engine == 2) {
        if (dtype != VAE2_DT_F32) return VAE2_ERR_ARG;
        return conv_dgrad_tf32(dy, w_packed_t, dx, G(g), accumulate, S(stream));
    }
    return conv_dgrad_simt(dy, reinterpret_cast<const float*>(w_packed_t), dx, dtype, G(g), accumulate, S(stream));
}
int vae2_conv2d_wgrad(const void* x, const void* dy, float* dw_packed, int dtype, const vae2_conv_geom* g, int engine,
                      vae2_stream_t stream) {
    if (engine != 0) return VAE2_ERR_UNSUPPORTED;
    return conv_wgrad_simt(x, dy, dw_packed, dtype, G(g), S(stream));
}
int vae2_bias_grad(const void* dy, float* dbias, int dtype, int64_t npix, int C, int ld, int accumulate,
                   vae2_stream_t stream) {
    return bias_grad(dy, dbias, dtype, npix, C, ld, accumulate, S(stream));
}
int vae2_conv2d_tc_supported(const vae2_conv_geom* g) { return conv_tc_supported(G(g)); }
int vae2_conv2d_tf32_supported(const vae2_conv_geom* g) { return conv_tf32_supported(G(g)); }
void vae2_conv2d_tf32_dims(const vae2_conv_geom* g, int* Nf, int* Kf, int* NfT, int* KfT) { conv_tf32_dims(G(g), Nf, Kf, NfT, KfT); }
int vae2_pack_weights_tf32(const vae2_tf32_pack_desc* d, int n, vae2_stream_t stream) {
    return pack_weights_tf32(reinterpret_cast<const Tf32PackDesc*>(d), n, S(stream));
}
long long vae2_conv2d_wgrad_tc_workspace(const vae2_conv_geom* g) { return conv_wgrad_tc_workspace(G(g)); }
int vae2_conv2d_wgrad_tc(const void* x, const void* dy, float* dw_packed, float* workspace, const vae2_conv_geom* g,
                         vae2_stream_t stream) {
    return conv_wgrad_tc(x, dy, dw_packed, workspace, G(g), S(stream));
}

long long vae2_conv2d_wgrad_f32x2_workspace(const vae2_conv_geom* g) { return conv_wgrad_f32x2_workspace(G(g)); }
int vae2_conv2d_wgrad_f32x2(const float* x, const float* dy, float* dw_packed, void* workspace, const vae2_conv_geom* g,
                            vae2_stream_t stream) {
    return conv_wgrad_f32x2(x, dy, dw_packed, workspace, G(g), S(stream));
}

int vae2_bn_max_partials(void) { return bn_stats_max_partials(); }
int vae2_bn_stats(const void* y, float* partials, int* n_partials, int dtype, int64_t npix, int Cp, int ld,
                  vae2_stream_t stream) {
    return bn_stats(y, partials, n_partials, dtype, npix, Cp, ld, S(stream));
}
int vae2_bn_merge(const float* partials, int n_partials, int Cp, float* merged, vae2_stream_t stream) {
    return bn_merge(partials, n_partials, Cp, merged, S(stream));
}
int vae2_bn_finalize(const float* partials, int n_partials, int C, int Cp, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps, float* mean,
                     float* invstd, float* scale, float* shift, vae2_stream_t stream) {
    return bn_finalize(partials, n_partials, C, Cp, gamma, beta, running_mean, running_var,
                       reinterpret_cast<long long*>(nbt), momentum, eps, mean, invstd, scale, shift, S(stream));
}
int vae2_bn_finalize_strided(const float* partials, int n_partials, int64_t part_stride, int C, int Cp, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, int64_t* nbt, float momentum,
                             float eps, float* mean, float* invstd, float* scale, float* shift, vae2_stream_t stream) {
    return bn_finalize(partials, n_partials, C, Cp, gamma, beta, running_mean, running_var,
                       reinterpret_cast<long long*>(nbt), momentum, eps, mean, invstd, scale, shift, S(stream), part_stride);
}
int vae2_bn_eval_coeffs(int C, int Cp, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, vae2_stream_t stream) {
    return bn_eval_coeffs(C, Cp, gamma, beta, running_mean, running_var, eps, scale, shift, S(stream));
}
int vae2_bn_apply(const void* y, const void* res, void* out, int dtype, int64_t npix, int Cp, int ld_y, int ld_res,
                  int ld_out, const float* scale, const float* shift, int relu, vae2_stream_t stream) {
    return bn_apply(y, res, out, dtype, npix, Cp, ld_y, ld_res, ld_out, scale, shift, relu, S(stream));
}
int vae2_bn_bwd_reduce(const void* g, const void* a, const void* y, float* partials, int* n_partials, int dtype,
                       int64_t npix, int Cp, int ld_g, int ld_a, int ld_y, const float* mean, const float* invstd, int relu,
                       vae2_stream_t stream) {
    return bn_bwd_reduce(g, a, y, partials, n_partials, dtype, npix, Cp, ld_g, ld_a, ld_y, mean, invstd, relu, S(stream));
}
int vae2_debug_bn_phase_times(uint64_t* device_buf6) { bn_debug_set_prof((unsigned long long*)device_buf6); return VAE2_OK; }
int vae2_bn_fwd_fused(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C, int Cp,
                      int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* mean,
                      float* invstd, float* scale, float* shift, int relu, vae2_stream_t stream) {
    return bn_fwd_fused(y, res, out, partials, dtype, npix, C, Cp, ld_y, ld_res, ld_out, gamma, beta, running_mean,
                        running_var, (long long*)num_batches_tracked, momentum, eps, mean, invstd, scale, shift, relu,
                        S(stream));
}
int vae2_bn_bwd_fused(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                      int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                      const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                      float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                      vae2_stream_t stream) {
    return bn_bwd_fused(g, a, y, dy, dres, partials, dtype, npix, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd,
                        scale, shift, dgamma, dbeta, accumulate_param, c1, c2, relu, acc_dy, acc_dres, S(stream));
}
int vae2_bn_fwd_fused_groups(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C,
                             int Cp, int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                             float eps, float* mean, float* invstd, float* scale, float* shift, int relu, int groups,
                             int stat_stride, vae2_stream_t stream) {
    return bn_fwd_fused_groups(y, res, out, partials, dtype, npix, C, Cp, ld_y, ld_res, ld_out, gamma, beta, running_mean,
                               running_var, (long long*)num_batches_tracked, momentum, eps, mean, invstd, scale, shift, relu,
                               groups, stat_stride, S(stream));
}
int vae2_bn_bwd_fused_groups(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                             int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                             const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                             float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                             int groups, int stat_stride, vae2_stream_t stream) {
    return bn_bwd_fused_groups(g, a, y, dy, dres, partials, dtype, npix, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd,
                               scale, shift, dgamma, dbeta, accumulate_param, c1, c2, relu, acc_dy, acc_dres, groups,
                               stat_stride, S(stream));
}
int vae2_bn_sync_fwd_stats(const void* y, float* partials, int dtype, int64_t npix, int C, int Cp, int ld_y, int groups,
                           float* msg, vae2_stream_t stream) {
    return bn_sync_fwd_stats(y, partials, dtype, npix, C, Cp, ld_y, groups, msg, S(stream));
}
int vae2_bn_sync_fwd_apply(const void* y, const void* res, void* out, int dtype, int64_t npix, int C, int Cp, int ld_y,
                           int ld_res, int ld_out, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* mean,
                           float* invstd, float* scale, float* shift, int relu, int groups, int stat_stride,
                           const float* gathered, int n_parts, int64_t part_stride, vae2_stream_t stream) {
    return bn_sync_fwd_apply(y, res, out, dtype, npix, C, Cp, ld_y, ld_res, ld_out, gamma, beta, running_mean, running_var,
                             (long long*)num_batches_tracked, momentum, eps, mean, invstd, scale, shift, relu, groups,
                             stat_stride, gathered, n_parts, part_stride, S(stream));
}
int vae2_bn_sync_bwd(int phase, const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                     int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean,
                     const float* invstd, const float* scale, const float* shift, float* dgamma, float* dbeta,
                     int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres, int groups,
                     int stat_stride, float* msg, const float* gsum, float inv_count, vae2_stream_t stream) {
    return bn_sync_bwd(phase, g, a, y, dy, dres, partials, dtype, npix, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd,
                       scale, shift, dgamma, dbeta, accumulate_param, c1, c2, relu, acc_dy, acc_dres, groups, stat_stride, msg,
                       gsum, inv_count, S(stream));
}
int vae2_ipc_alloc(int64_t bytes, void** ptr, void* handle64) { return ipc_alloc(bytes, ptr, handle64); }
int vae2_ipc_open(const void* handle64, void** ptr) { return ipc_open(handle64, ptr); }
int vae2_ipc_close(void* ptr) { return ipc_close(ptr); }
int vae2_ipc_free(void* ptr) { return ipc_free(ptr); }
int vae2_bn_peer_setup(int world, int rank, void* const* bases, uint32_t* seq, int* err) {
    return bn_peer_setup(world, rank, bases, seq, err);
}
int64_t vae2_bn_peer_slot_words(int world, int groups, int Cp, int backward) {
    return bn_peer_slot_words(world, groups, Cp, backward);
}
int vae2_bn_fwd_fused_peer(const void* y, const void* res, void* out, float* partials, int dtype, int64_t npix, int C,
                           int Cp, int ld_y, int ld_res, int ld_out, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                           float eps, float* mean, float* invstd, float* scale, float* shift, int relu, int groups,
                           int stat_stride, int64_t slot_word, int seq_index, vae2_stream_t stream) {
    return bn_fwd_fused_peer(y, res, out, partials, dtype, npix, C, Cp, ld_y, ld_res, ld_out, gamma, beta, running_mean,
                             running_var, (long long*)num_batches_tracked, momentum, eps, mean, invstd, scale, shift, relu,
                             groups, stat_stride, slot_word, seq_index, S(stream));
}
int vae2_bn_bwd_fused_peer(const void* g, const void* a, const void* y, void* dy, void* dres, float* partials, int dtype,
                           int64_t npix, int C, int Cp, int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres,
                           const float* mean, const float* invstd, const float* scale, const float* shift, float* dgamma,
                           float* dbeta, int accumulate_param, float* c1, float* c2, int relu, int acc_dy, int acc_dres,
                           int groups, int stat_stride, float inv_count, int64_t slot_word, int seq_index,
                           vae2_stream_t stream) {
    return bn_bwd_fused_peer(g, a, y, dy, dres, partials, dtype, npix, C, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd,
                             scale, shift, dgamma, dbeta, accumulate_param, c1, c2, relu, acc_dy, acc_dres, groups,
                             stat_stride, inv_count, slot_word, seq_index, S(stream));
}
int vae2_bn_bwd_finalize(const float* partials, int n_partials, int C, int Cp, float* sums, vae2_stream_t stream) {
    return bn_bwd_finalize(partials, n_partials, C, Cp, sums, S(stream));
}
int vae2_bn_bwd_coeffs(const float* sums_global, int C, int Cp, float inv_count, float* dgamma, float* dbeta,
                       int accumulate_param, const float* sums_local, float* c1, float* c2, vae2_stream_t stream) {
    return bn_bwd_coeffs(sums_global, C, Cp, inv_count, dgamma, dbeta, accumulate_param, sums_local, c1, c2, S(stream));
}
int vae2_bn_bwd_elemt(const void* g, const void* a, const void* y, void* dy, void* dres, int dtype, int64_t npix, int Cp,
                      int ld_g, int ld_a, int ld_y, int ld_dy, int ld_dres, const float* mean, const float* invstd,
                      const float* scale, const float* c1, const float* c2, int relu, int acc_dy, int acc_dres,
                      vae2_stream_t stream) {
    return bn_bwd_elemt(g, a, y, dy, dres, dtype, npix, Cp, ld_g, ld_a, ld_y, ld_dy, ld_dres, mean, invstd, scale, c1, c2,
                        relu, acc_dy, acc_dres, S(stream));
}

int vae2_fuse_sum(const vae2_fuse_src* srcs, int nsrc, void* out, int dtype, int B, int H, int W, int Cp, int ld_out,
                  int relu, vae2_stream_t stream) {
    return fuse_sum(reinterpret_cast<const FuseSrc*>(srcs), nsrc, out, dtype, B, H, W, Cp, ld_out, relu, S(stream));
}
int vae2_fuse_bwd_same(const void* g, const void* out, const vae2_fuse_dst* dsts, int ndst, int dtype, int64_t npix, int Cp,
                       int ld_g, int ld_out, int relu, vae2_stream_t stream) {
    return fuse_bwd_same(g, out, reinterpret_cast<const FuseDst*>(dsts), ndst, dtype, npix, Cp, ld_g, ld_out, relu, S(stream));
}
int vae2_fuse_bwd_up(const void* g, const void* out, void* gsrc, int dtype, int B, int H, int W, int Hs, int Ws, int Cp,
                     int ld_g, int ld_out, int ld_gsrc, int relu, int accumulate, vae2_stream_t stream) {
    return fuse_bwd_up(g, out, gsrc, dtype, B, H, W, Hs, Ws, Cp, ld_g, ld_out, ld_gsrc, relu, accumulate, S(stream));
}

int vae2_elbo_acc_floats(void) { return elbo_acc_floats(); }
int vae2_elbo_terms(const vae2_elbo_seg* segs_dev, int nseg, float* acc, int nslots, int32_t* nonfinite,
                    vae2_stream_t stream) {
    return elbo_terms(reinterpret_cast<const ElboSeg*>(segs_dev), nseg, acc, nslots, nonfinite, S(stream));
}
int vae2_elbo_terms_bwd(const vae2_elbo_bwd_seg* segs_dev, int nseg, vae2_stream_t stream) {
    return elbo_terms_bwd(reinterpret_cast<const ElboBwdSeg*>(segs_dev), nseg, S(stream));
}

int vae2_clip_u8_to_nchw(const uint8_t* frames, float* dst, int B, int L, int H, int W, vae2_stream_t stream) {
    return clip_u8_to_nchw(frames, dst, B, L, H, W, S(stream));
}
int vae2_to_image(const float* x, float* im, int64_t n, int HW, vae2_stream_t stream) { return to_image(x, im, n, HW, S(stream)); }
int vae2_frame_metrics(const float* pred_im, const float* gt_im, double* out, int R, int F, int Bg, int frame_elems,
                       vae2_stream_t stream) {
    return frame_metrics(pred_im, gt_im, out, R, F, Bg, frame_elems, S(stream));
}
int vae2_ssim_level(const float* X, const float* Y, double* out, int N, int Ny, int H, int W, float data_range,
                    vae2_stream_t stream) {
    return ssim_level(X, Y, out, N, Ny, H, W, data_range, S(stream));
}
int vae2_avgpool2(const float* x, float* y, int N, int H, int W, vae2_stream_t stream) { return avgpool2(x, y, N, H, W, S(stream)); }

int vae2_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, const int64_t* step_dev, float grad_scale, vae2_stream_t stream) {
    return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, reinterpret_cast<const long long*>(step_dev),
                     grad_scale, S(stream));
}

}  // extern "C"
