"""Device input pipeline: uint8 frames -> normalised clips on the GPU.

The reference prepares every clip on the host in numpy (lib/datasets/cityscapes.py:300-326: float32 frames, /255,
ImageNet mean/std, HWC -> CHW, 3 clips x 3 frames stacked along channels) and ships 3 x [B, 9, H, W] float32 tensors to
the device (lib/core/function.py:487-489): 12 bytes per pixel-channel over PCIe / C2C per step.  Here the host side only
stacks the decoded uint8 RGB frames; one kernel (csrc/metrics.cu: clip_u8_to_nchw) does the arithmetic and the layout
change on the device -- 1 byte per pixel-channel crosses the bus -- and a two-deep pinned staging ring overlaps the copy
of batch i+1 with the step on batch i.  Values are those of the reference's float32 numpy expression to within 1 ulp.
"""
import os
import sys

import torch

_LIB = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _LIB not in sys.path:
    sys.path.insert(0, _LIB)
from _engine_loader import engine  # noqa: E402

_E = engine()
_N = _E.native


def clips_from_u8(frames, clip_num=3):
    """frames: uint8 CUDA tensor [B, clip_num*L, H, W, 3] (RGB frames of a sample in temporal order) ->
    list of clip_num tensors [B, 3L, H, W] float32, normalised like cityscapes.py:311-316."""
    if frames.device.type != "cuda" or frames.dtype != torch.uint8:
        raise RuntimeError("vae2_b200: clips_from_u8 expects a uint8 CUDA tensor; there is no CPU fallback")
    B, T, H, W, C_ = frames.shape
    assert C_ == 3 and T % clip_num == 0
    out = torch.empty(B, T * 3, H, W, dtype=torch.float32, device=frames.device)
    _N.call.vae2_clip_u8_to_nchw(frames.contiguous().data_ptr(), out.data_ptr(), B, T, H, W,
                                 torch.cuda.current_stream(frames.device).cuda_stream)
    L3 = (T // clip_num) * 3
    return [out[:, i * L3:(i + 1) * L3] for i in range(clip_num)]


class DeviceClipLoader:
    """Wraps an iterable of uint8 host batches [B, 9, H, W, 3] (e.g. a DataLoader over decoded frames): pinned
    double-buffered staging, asynchronous H2D on a side stream, normalisation on the device.  Yields (xt, x2t, x3t)."""

    def __init__(self, batches, device, clip_num=3):
        self.batches, self.device, self.clip_num = batches, torch.device(device), clip_num
        self.copy_stream = torch.cuda.Stream(self.device)
        self.bytes_per_batch = 0

    def __iter__(self):
        ring, pending = [None, None], None
        it = iter(self.batches)
        i = 0

        def stage(host, slot):
            if ring[slot] is None or ring[slot][0].shape != host.shape:
                ring[slot] = (torch.empty(host.shape, dtype=torch.uint8).pin_memory(),
                              torch.empty(host.shape, dtype=torch.uint8, device=self.device), torch.cuda.Event(),
                              torch.cuda.Event())
            pin, dev, ev, consumed = ring[slot]
            ev.synchronize()                       # the device copy that last used this pinned buffer has finished
            pin.copy_(host)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(consumed)   # the normalise kernel that last read this device buffer is done
                dev.copy_(pin, non_blocking=True)
                ev.record(self.copy_stream)
            self.bytes_per_batch = host.numel()
            return dev, ev, consumed

        try:
            pending = stage(next(it), 0)
        except StopIteration:
            return
        while pending is not None:
            dev, ev, consumed = pending
            i += 1
            try:
                nxt = stage(next(it), i % 2)
            except StopIteration:
                nxt = None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            clips = tuple(clips_from_u8(dev, self.clip_num))
            consumed.record(cur)
            yield clips
            pending = nxt
