"""Segmentation HRNet entry point kept for tools/test.py (reference lib/models/seg_hrnet.py:476-480).

SURVEY.md §8 marks the legacy segmentation path out of scope except that ``tools/test.py``
must be able to import ``models.seg_hrnet``; its blocks are the ones in ``models.enc_hrnet``.
"""
from .enc_hrnet import BasicBlock, Bottleneck, HighResolutionModule, blocks_dict  # noqa: F401


def get_seg_model(cfg, **kwargs):
    raise NotImplementedError(
        "vae2_b200: the HRNetV2 segmentation network (stride-2 stem, single seg head) is outside the "
        "VAE^2 hot path built here; use models.enc_hrnet.get_* for the VAE^2 networks")
