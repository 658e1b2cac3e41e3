"""K-sample inference driver and device-side metrics  (SURVEY.md §8 f1 / f4).

The reference's ``inference`` (lib/core/function.py:55-440) draws NUM_SAMPLES predictions per context clip by running the
WHOLE ``FullModel_encdec`` once per draw (:124-146: posterior net, encoder + decoders, four discriminator passes), then
un-normalises every prediction on the host and scores it with numpy / pytorch_msssim on the CPU (:238-316).

Here one call does all K draws of a batch of clips:
  * prior sampling needs neither the posterior net nor the discriminators (their outputs only feed loss terms the
    driver discards), so they do not run;
  * the encoder trunk up to ``transition3`` does not depend on z: it runs once per clip, its features are replicated K
    times (engine TileOp) and everything downstream runs once at batch K*B (``HighResolutionNetED.sample_k``);
  * eval-mode BN has no batch statistics, so stacking the draws is exactly K independent calls;
  * un-normalisation, L1 / PSNR and SSIM / MS-SSIM run on the device (csrc/metrics.cu) -- no prediction leaves the GPU
    unless the caller asks for it.

RNG: eps is drawn with ``torch.randn`` in the reference's order -- per draw the four z maps (utils.py:88-90), then the
encoder's code (enc_hrnet.py:456) -- so a seeded run reproduces the reference's draws.
"""
import math
import os
import sys

import torch

_LIB = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _LIB not in sys.path:
    sys.path.insert(0, _LIB)
from _engine_loader import engine  # noqa: E402

_E = engine()
_N = _E.native


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(t):
    if t.device.type != "cuda":
        raise RuntimeError("vae2_b200: device-side metrics run on CUDA tensors only; there is no CPU fallback")


def to_image(x):
    """``_to_image(x, is_uint8=False)`` (function.py:87-98) for nchw tensors whose channels are RGB triplets:
    (x*std + mean)*255 clipped to [0, 255], same layout."""
    _need_cuda(x)
    x = x.contiguous().float()
    assert x.dim() == 4 and x.shape[1] % 3 == 0
    im = torch.empty_like(x)
    _N.call.vae2_to_image(x.data_ptr(), im.data_ptr(), x.numel(), x.shape[2] * x.shape[3], _st(x.device))
    return im


def frame_metrics(pred_im, gt_im):
    """Per (row, frame) mean |pred - gt| and PSNR (function.py:262-263, criterion.py:106-116) of image-space tensors
    [R, 3F, H, W] vs [B, 3F, H, W] (row r is compared with gt row r % B).  Returns (recon [R, F], psnr [R, F]), float64."""
    _need_cuda(pred_im)
    R, C3, H, W = pred_im.shape
    F_, Bg = C3 // 3, gt_im.shape[0]
    assert gt_im.shape[1:] == pred_im.shape[1:] and R % Bg == 0
    out = torch.empty(R, F_, 2, dtype=torch.float64, device=pred_im.device)
    _N.call.vae2_frame_metrics(pred_im.contiguous().data_ptr(), gt_im.contiguous().data_ptr(), out.data_ptr(), R, F_, Bg,
                               3 * H * W, _st(pred_im.device))
    n = 3.0 * H * W
    recon = out[..., 0] / n
    psnr = 20.0 * torch.log10(255.0 / torch.sqrt(out[..., 1] / n))
    return recon, psnr


def _ssim_level(X, Y, data_range):
    """planes X [N, H, W] vs Y [Ny, H, W] -> (ssim_per_plane [N], cs_per_plane [N]) (means over valid positions)."""
    N_, H, W = X.shape
    out = torch.empty(N_, 2, dtype=torch.float64, device=X.device)
    _N.call.vae2_ssim_level(X.data_ptr(), Y.data_ptr(), out.data_ptr(), N_, Y.shape[0], H, W, float(data_range), _st(X.device))
    cnt = float((H - 10) * (W - 10))
    return out[:, 0] / cnt, out[:, 1] / cnt


def _pool(X):
    N_, H, W = X.shape
    Ho, Wo = (H + 2 * (H % 2) - 2) // 2 + 1, (W + 2 * (W % 2) - 2) // 2 + 1
    Y = torch.empty(N_, Ho, Wo, dtype=torch.float32, device=X.device)
    _N.call.vae2_avgpool2(X.data_ptr(), Y.data_ptr(), N_, H, W, _st(X.device))
    return Y


def ssim(X, Y, data_range=255, size_average=True):
    """pytorch_msssim.ssim (11-tap Gaussian, sigma 1.5, K=(0.01, 0.03)) for [N, C, H, W] float images; Y may hold fewer
    images than X (image n is compared with n % len(Y)).  size_average=False returns one value per image."""
    _need_cuda(X)
    N_, C_, H, W = X.shape
    s, _ = _ssim_level(X.contiguous().float().view(N_ * C_, H, W), Y.contiguous().float().view(-1, H, W), data_range)
    per_img = s.view(N_, C_).mean(1)
    return per_img.mean() if size_average else per_img


def ms_ssim(X, Y, data_range=255, size_average=True, weights=(1.0 / 3, 1.0 / 3, 1.0 / 3)):
    """pytorch_msssim.ms_ssim with the weights the reference passes (function.py:25): per level SSIM / contrast
    sensitivity, 2x2 average pooling (padding = size % 2) between levels, product of relu(cs_l)^w_l and relu(ssim_L)^w_L."""
    _need_cuda(X)
    N_, C_, H, W = X.shape
    assert min(H, W) > (11 - 1) * 2 ** (len(weights) - 1), "image too small for this many MS-SSIM levels"
    x, y = X.contiguous().float().view(N_ * C_, H, W), Y.contiguous().float().view(-1, H, W)
    vals = []
    for i in range(len(weights)):
        s, cs = _ssim_level(x, y, data_range)
        if i < len(weights) - 1:
            vals.append(torch.relu(cs))
            x, y = _pool(x), _pool(y)
    vals.append(torch.relu(s))
    w = torch.tensor(weights, dtype=torch.float64, device=X.device).view(-1, 1)
    per_plane = torch.prod(torch.stack(vals, 0) ** w, dim=0)
    per_img = per_plane.view(N_, C_).mean(1)
    return per_img.mean() if size_average else per_img


class KSampleInference:
    """``KSampleInference(model_encdec, K)(xt, x2t, x3t)`` -> dict with the K predictions of every clip and their
    per-frame scores.  ``model_encdec`` is the ``FullModel_encdec`` wrapper (or its encoder/decoder net) in eval mode."""

    def __init__(self, model_encdec, K=16, with_ssim=True, keep_predictions=True):
        self.net = getattr(model_encdec, "encdec_model", model_encdec)
        self.K, self.with_ssim, self.keep = int(K), with_ssim, keep_predictions

    def draw(self, B, H, W, device):
        """eps in the reference's order: for each draw, z maps of the 4 branches, then the code."""
        Z = self.net.z_dim
        sizes, (h, w) = [], (H, W)
        for _ in range(4):
            sizes.append((h, w))
            h, w = (h + 1) // 2, (w + 1) // 2
        zs, codes = [[] for _ in range(4)], []
        for _ in range(self.K):
            for i, (h, w) in enumerate(sizes):
                zs[i].append(torch.randn(B, Z, h, w, device=device))
            codes.append(torch.randn(B, Z, 1, 1, device=device))
        return [torch.cat(z, 0) for z in zs], torch.cat(codes, 0)

    @torch.no_grad()
    def __call__(self, xt, x2t, x3t, eps=None):
        if self.net.training:
            raise RuntimeError("vae2_b200: KSampleInference needs model.eval() (the reference's inference() sets it, function.py:60)")
        B, _, H, W = xt.shape
        z, code = eps if eps is not None else self.draw(B, H, W, xt.device)
        x1p, x2p, x3p = self.net.sample_k(xt, z, code)
        out = {"K": self.K}
        gt2, gt3 = to_image(x2t), to_image(x3t)
        for name, pred, gt in (("x2t", x2p, gt2), ("x3t", x3p, gt3)):
            im = to_image(pred)
            recon, psnr = frame_metrics(im, gt)
            out[name + "_recon"], out[name + "_psnr"] = recon.view(self.K, B, -1), psnr.view(self.K, B, -1)
            if self.with_ssim:
                F_ = im.shape[1] // 3
                fr, gf = im.view(-1, 3, H, W), gt.view(-1, 3, H, W)        # one image per (row, frame)
                # frame j of prediction row r pairs with frame j of gt row r % B: index r*F + j -> ((r % B)*F + j)
                gidx = (torch.arange(fr.shape[0], device=im.device) // F_ % B) * F_ + torch.arange(fr.shape[0], device=im.device) % F_
                gsel = gf[gidx]
                out[name + "_ssim"] = ssim(fr, gsel, 255, size_average=False).view(self.K, B, F_)
                if min(H, W) > (11 - 1) * 2 ** 2:          # pytorch_msssim's size requirement for the 3 levels used
                    out[name + "_msssim"] = ms_ssim(fr, gsel, 255, size_average=False).view(self.K, B, F_)
        if self.keep:
            out["xt_predict"], out["x2t_predict"], out["x3t_predict"] = (t.view(self.K, B, *t.shape[1:]) for t in (x1p, x2p, x3p))
        return out
