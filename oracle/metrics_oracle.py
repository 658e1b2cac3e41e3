"""CPU oracle for the callers either side of the path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with numpy / torch CPU ops:
  * ``input_transform`` + ``__getitem__`` of the dataset (lib/datasets/cityscapes.py:307-326): frames -> normalised clips;
  * ``_to_image(x, is_uint8=False)`` (lib/core/function.py:87-98), the per-frame recon loss (:262) and ``PSNR``
    (lib/core/criterion.py:106-116);
  * ``ssim`` / ``ms_ssim`` of the third-party package **pytorch_msssim** as the reference calls them
    (function.py:24-25, 244-261: data_range=255, size_average=True, ms_ssim weights [1/3]*3).

pytorch_msssim is NOT in /root/reference nor in this image, and the reference's requirements.txt lists it without a
version: **parity unpinned** for the two SSIM functions.  The algorithm restated here is the package's published one
(v0.2.x ``_ssim`` / ``ms_ssim``): 11-tap Gaussian window (sigma 1.5) applied separably per channel with VALID
convolution, K = (0.01, 0.03), C_i = (K_i * data_range)^2, cs = (2 s12 + C2)/(s1 + s2 + C2),
ssim = (2 mu1 mu2 + C1)/(mu1^2 + mu2^2 + C1) * cs, spatial mean per channel; MS-SSIM: per level relu(cs), 2x2 average
pooling with padding = size % 2 between levels, product of level values raised to the weights, mean over channels.
Closed-form anchors (tests/test_gpu_metrics.py): ssim(X, X) = 1; for constant images a, b:
ssim = (2ab + C1)/(a^2 + b^2 + C1); both pin the window normalisation, C1, C2 and data_range.
"""
import numpy as np
import torch
import torch.nn.functional as F

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)   # cityscapes.py:58-59 (ImageNet statistics)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def clips_from_frames(frames_u8, clip_num=3):
    """cityscapes.py:307-326 for one sample: list of T uint8 frames [H, W, 3] -> clip_num arrays [3L, H, W] float32."""
    T = len(frames_u8)
    seq = np.concatenate([np.asarray(f, dtype=np.float32) for f in frames_u8], axis=-1)
    seq = seq / 255.0
    seq -= np.tile(MEAN, T)          # self.mean * self.clip_length * self.clip_num (list repetition)
    seq /= np.tile(STD, T)
    seq = np.transpose(seq, (2, 0, 1))
    L3 = (T // clip_num) * 3
    return [seq[i * L3:(i + 1) * L3].copy() for i in range(clip_num)]


def to_image(x_chw):
    """function.py:87-98 with is_uint8=False; x_chw [3, H, W] float32 -> [H, W, 3] float32 in [0, 255]."""
    x = np.transpose(np.array(x_chw, dtype=np.float32), (1, 2, 0)).copy()
    x *= STD
    x += MEAN
    x *= 255.0
    np.clip(x, 0, 255, out=x)
    return x


def recon_and_psnr(im, im_gt):
    """function.py:262-263: np.mean(|im - im_gt|) and criterion.py:113-116."""
    recon = float(np.mean(np.abs(im - im_gt)))
    mse = torch.mean((torch.from_numpy(im) - torch.from_numpy(im_gt)) ** 2)
    return recon, float(20 * torch.log10(255.0 / torch.sqrt(mse)))


def _window(size=11, sigma=1.5):
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filt(x, win):
    C_ = x.shape[1]
    k = win.to(x.dtype)
    x = F.conv2d(x, k.view(1, 1, -1, 1).repeat(C_, 1, 1, 1), groups=C_)
    return F.conv2d(x, k.view(1, 1, 1, -1).repeat(C_, 1, 1, 1), groups=C_)


def _ssim(X, Y, data_range, win):
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _filt(X, win), _filt(Y, win)
    s1 = _filt(X * X, win) - mu1 * mu1
    s2 = _filt(Y * Y, win) - mu2 * mu2
    s12 = _filt(X * Y, win) - mu1 * mu2
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ssim(X, Y, data_range=255.0):
    """[N, C, H, W] -> per-image value (mean over channels)."""
    s, _ = _ssim(X, Y, data_range, _window())
    return s.mean(1)


def ms_ssim(X, Y, data_range=255.0, weights=(1.0 / 3, 1.0 / 3, 1.0 / 3)):
    win = _window()
    w = torch.tensor(weights, dtype=X.dtype)
    mcs = []
    for i in range(len(weights)):
        s, cs = _ssim(X, Y, data_range, win)
        if i < len(weights) - 1:
            mcs.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X, Y = F.avg_pool2d(X, 2, padding=pad), F.avg_pool2d(Y, 2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(s)], 0)
    return torch.prod(vals ** w.view(-1, 1, 1), dim=0).mean(1)
