"""Loss primitives of the VAE^2 ELBO on the fused CUDA kernel  --  drop-in for the reference's
lib/core/criterion.py:61-103 (``L1Loss``, ``KLLoss``, ``lsgan_adversarial_loss``; same call
keywords ``predict=/target=``, ``mu=/logvar=``, ``sample=/mode=``, same value: sum / batch).

Each module is one launch of the table-driven ELBO kernel (engine/elbo.py, csrc/elbo.cu);
``utils.utils.FullModel_encdec`` batches all terms of a step into two launches instead.
The segmentation losses of the reference (CrossEntropy/OHEM, :11-58) are outside the VAE^2 path and fall through
to the reference's own module (see _fallthrough.py).
"""
import os
import sys

import torch
import torch.nn as nn

_LIB = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _LIB not in sys.path:
    sys.path.insert(0, _LIB)
from _engine_loader import engine  # noqa: E402

import _fallthrough  # noqa: E402

_E = engine()

# legacy segmentation losses (reference :11-58) are outside the VAE^2 path: taken from the reference's own module
# when its lib/ is on sys.path behind this tree, so that ``from core.criterion import *`` (tools/train.py:32) sees them
_ref = _fallthrough.reference_module("core.criterion")
if _ref is not None:
    CrossEntropy, OhemCrossEntropy = _ref.CrossEntropy, _ref.OhemCrossEntropy


class L1Loss(nn.Module):
    """sum |predict - target| / B   (reference :61-69)"""

    def forward(self, predict, target):
        spec = [dict(kind=0, slot=0, a=0, b=1, scale=1.0 / predict.shape[0], name="predict")]
        return _E.elbo_terms(spec, 1, [predict, target.detach()])[0][0]


class KLLoss(nn.Module):
    """sum 0.5 (mu^2 + e^logvar - logvar - 1) / B, over a tensor or a list of maps (reference :72-87)"""

    def forward(self, mu, logvar):
        if not isinstance(mu, list):
            mu, logvar = [mu], [logvar]
        assert isinstance(logvar, list) and len(mu) == len(logvar)
        tensors, spec = [], []
        for m, v in zip(mu, logvar):
            m4 = m if m.dim() == 4 else m.reshape(m.shape[0], -1, 1, 1)
            v4 = v if v.dim() == 4 else v.reshape(v.shape[0], -1, 1, 1)
            tensors.append(torch.cat([m4, v4], 1))
            spec.append(dict(kind=1, slot=0, a=None, b=len(tensors) - 1, scale=1.0 / m.shape[0], name="kl"))
        return _E.elbo_terms(spec, 1, tensors)[0][0]


class lsgan_adversarial_loss(nn.Module):
    """sum (sample - 1)^2 / B for 'real', sum sample^2 / B for 'fake'   (reference :90-103)"""

    def forward(self, sample, mode):
        assert mode in ["real", "fake"]
        spec = [dict(kind=2, slot=0, a=0, b=None, scale=1.0 / sample.shape[0],
                     target=1.0 if mode == "real" else 0.0, name="sample")]
        return _E.elbo_terms(spec, 1, [sample])[0][0]


class PSNR:
    """Peak signal-to-noise ratio for images in [0, 255] (reference :106-116; evaluation metric)."""

    def __init__(self):
        self.name = "PSNR"

    @staticmethod
    def __call__(img1, img2):
        mse = torch.mean((img1 - img2) ** 2)
        return 20 * torch.log10(255.0 / torch.sqrt(mse))
