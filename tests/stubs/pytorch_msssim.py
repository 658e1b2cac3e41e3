"""Stub of pytorch_msssim for the drop-in tests: function.py:24 imports ssim / ms_ssim at module level; the training
loop under test never calls them."""


def ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim stub")


def ms_ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim stub")
