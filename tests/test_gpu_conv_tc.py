"""GPU: the tcgen05/TMEM/TMA convolution (engine 1) against torch CPU conv2d on bf16-rounded data."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from helpers import E, O, rel_err
from test_gpu_kernels import dev, st, pad, to_act, from_act, table

N = E.native
pytestmark = pytest.mark.gpu

TC_CASES = [  # (B, Cin, Cout, H, W, k, bias, stride)
    (1, 64, 64, 8, 128, 3, False, 1),     # one full-width strip per tile, KC=64 (SWIZZLE_128B)
    (1, 64, 64, 32, 64, 3, False, 1),     # TW=64, TH=2
    (2, 18, 18, 17, 23, 3, False, 1),     # odd size, Cin_p=32 (SWIZZLE_64B), patches overhang the image
    (1, 36, 36, 16, 32, 3, False, 1),     # Cin_p=48 -> KC=16 (SWIZZLE_32B), 3 chunks per tap
    (1, 64, 256, 9, 11, 1, False, 1),     # 1x1, N=256
    (1, 256, 64, 12, 40, 1, False, 1),    # 1x1, K=256 (4 chunks)
    (2, 270, 270, 16, 32, 1, True, 1),    # head conv: Cp=272, two N tiles of 144, bias
    (1, 144, 144, 8, 8, 3, False, 1),     # low-res branch, TW=8 TH=16
    (1, 72, 72, 33, 47, 3, False, 1),
    (3, 9, 64, 20, 36, 3, False, 1),      # stem
    (2, 18, 36, 16, 32, 3, False, 2),     # stride 2 (fuse-down / new-branch transitions), even size
    (1, 36, 72, 33, 47, 3, False, 2),     # stride 2, odd size (Ho = ceil(H/2))
    (1, 256, 36, 24, 40, 3, False, 2),    # transition1 new branch
    (2, 72, 144, 9, 12, 3, False, 2),
    (4, 18, 18, 64, 128, 3, False, 1),    # halo weight gradient: several patches per split-K range (stage ring wraps)
    (2, 30, 150, 19, 9, 3, False, 1),     # halo weight gradient: Cin_p=32, N=160 (3 x 160 TMEM columns), ragged patches
    (1, 64, 72, 40, 24, 3, False, 1),     # halo weight gradient: Cin_p=64, two M=128 instructions per kernel row, N=80
    (2, 256, 18, 19, 21, 3, False, 1),    # halo weight gradient: four 64-lane channel chunks handled by different CTAs (transition layers)
]


def pack_bf16(w, Cin_p, Cout_p):
    Cout, Cin, k, _ = w.shape
    wd = w.contiguous().to(dev())
    n = k * k * Cin_p * Cout_p
    wq = torch.zeros(n, dtype=torch.bfloat16, device=dev())
    wqT = torch.zeros(n, dtype=torch.bfloat16, device=dev())
    d = (N.PackDesc * 1)()
    d[0] = N.PackDesc(w=wd.data_ptr(), wq=wq.data_ptr(), wqT=wqT.data_ptr(), Cout=Cout, Cin=Cin, k=k, Cin_p=Cin_p, Cout_p=Cout_p)
    t = table(d)
    N.call.vae2_pack_weights(t.data_ptr(), 1, st())
    torch.cuda.synchronize()
    return wq, wqT


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fwd_dgrad(case):
    B, Cin, Cout, H, W, k, use_bias, stride = case
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    tag = "tc%s" % (case,)
    x = O.det_normal(tag + "x", (B, Cin, H, W)).bfloat16().float()
    w = O.det_normal(tag + "w", (Cout, Cin, k, k), (2.0 / (Cin * k * k)) ** 0.5).bfloat16().float()
    bias = O.det_normal(tag + "b", (Cout,), 0.1) if use_bias else None
    xr = x.clone().requires_grad_(True)
    yr = F.conv2d(xr, w, bias, stride=stride, padding=k // 2)
    gy = O.det_normal(tag + "gy", tuple(yr.shape)).bfloat16().float()
    yr.backward(gy)

    xa, Cin_p = to_act(x, "bf16", pad(Cin, 16))
    Cout_p = pad(Cout, 16)
    wq, wqT = pack_bf16(w, Cin_p, Cout_p)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin_p, ldx=Cin_p, Ho=Ho, Wo=Wo, Cout_p=Cout_p, ldy=Cout_p, k=k, stride=stride, pad=k // 2)
    assert N.lib().vae2_conv2d_tc_supported(C.byref(g)) == 1
    ya = torch.zeros(B * Ho * Wo * Cout_p, dtype=torch.bfloat16, device=dev())
    bp = None
    if bias is not None:
        bp = torch.zeros(Cout_p, dtype=torch.float32, device=dev())
        bp[:Cout] = bias.to(dev())
    N.call.vae2_conv2d_fwd(xa.data_ptr(), wq.data_ptr(), bp.data_ptr() if bp is not None else None, ya.data_ptr(), 1,
                           C.byref(g), 1, st())
    torch.cuda.synchronize()
    y = from_act(ya, "bf16", B, Cout, Ho, Wo, Cout_p)
    assert rel_err(y, yr.detach()) < 6e-3, "tc fwd rel err %.3e" % rel_err(y, yr.detach())
    assert float(ya.view(-1, Cout_p)[:, Cout:].float().abs().sum()) == 0.0

    gya, _ = to_act(gy, "bf16", Cout_p)
    dxa = torch.zeros_like(xa)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wqT.data_ptr(), dxa.data_ptr(), 1, C.byref(g), 0, 1, st())
    torch.cuda.synchronize()
    dx = from_act(dxa, "bf16", B, Cin, H, W, Cin_p)
    assert rel_err(dx, xr.grad) < 6e-3, "tc dgrad rel err %.3e" % rel_err(dx, xr.grad)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wqT.data_ptr(), dxa.data_ptr(), 1, C.byref(g), 1, 1, st())
    torch.cuda.synchronize()
    assert rel_err(from_act(dxa, "bf16", B, Cin, H, W, Cin_p), 2 * xr.grad) < 1e-2, "tc dgrad accumulate"

    # weight gradient on tensor cores (split-K over pixels + workspace reduce)
    need = N.lib().vae2_conv2d_wgrad_tc_workspace(C.byref(g))
    if need > 0:
        wr = w.clone().requires_grad_(True)
        F.conv2d(x, wr, None, stride=stride, padding=k // 2).backward(gy)
        ws = torch.zeros(need, dtype=torch.float32, device=dev())
        dwp = torch.full((k * k * Cin_p * Cout_p,), 7.0, dtype=torch.float32, device=dev())   # must be overwritten
        N.call.vae2_conv2d_wgrad_tc(xa.data_ptr(), gya.data_ptr(), dwp.data_ptr(), ws.data_ptr(), C.byref(g), st())
        dw = torch.zeros_like(w).to(dev())
        d = (N.PackDesc * 1)()
        d[0] = N.PackDesc(w=dw.data_ptr(), wp=dwp.data_ptr(), Cout=Cout, Cin=Cin, k=k, Cin_p=Cin_p, Cout_p=Cout_p)
        t = table(d)
        N.call.vae2_unpack_wgrad(t.data_ptr(), 1, 0, st())
        torch.cuda.synchronize()
        assert rel_err(dw.cpu(), wr.grad) < 6e-3, "tc wgrad rel err %.3e" % rel_err(dw.cpu(), wr.grad)
    else:
        raise AssertionError("wgrad_tc unexpectedly unsupported")


def test_conv_tc_many_tiles_persistent():
    """More tiles than SMs (persistent loop, both TMEM accumulator stages, stage ring wrap-around)."""
    B, Cin, Cout, H, W, k = 2, 64, 64, 128, 256, 3
    x = O.det_normal("tcbig:x", (B, Cin, H, W)).bfloat16().float()
    w = O.det_normal("tcbig:w", (Cout, Cin, k, k), 0.05).bfloat16().float()
    yr = F.conv2d(x, w, None, 1, 1)
    xa, Cin_p = to_act(x, "bf16", 64)
    wq, _ = pack_bf16(w, 64, 64)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=64, ldx=64, Ho=H, Wo=W, Cout_p=64, ldy=64, k=k, stride=1, pad=1)
    ya = torch.zeros(B * H * W * 64, dtype=torch.bfloat16, device=dev())
    N.call.vae2_conv2d_fwd(xa.data_ptr(), wq.data_ptr(), None, ya.data_ptr(), 1, C.byref(g), 1, st())
    torch.cuda.synchronize()
    assert rel_err(from_act(ya, "bf16", B, Cout, H, W, 64), yr) < 6e-3


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_f32x3_fwd_dgrad(case):
    """The tensor-core convolution of the fp32 path (engine 2: exact 3-way bf16 split, 6 products, leading product and
    corrections in separate TMEM accumulators) against a float64 convolution.  Operands and products are exact; what remains
    is the tensor core's truncating fp32 accumulation of the leading product (round 1, one accumulator: 1e-6 .. 1.5e-5 per
    conv; now 7e-8 .. 2.3e-6, logged per case next to torch's own fp32 conv error)."""
    B, Cin, Cout, H, W, k, use_bias, stride = case
    tag = "t32%s" % (case,)
    x = O.det_normal(tag + "x", (B, Cin, H, W))
    w = O.det_normal(tag + "w", (Cout, Cin, k, k), (2.0 / (Cin * k * k)) ** 0.5)
    bias = O.det_normal(tag + "b", (Cout,), 0.1) if use_bias else None
    xr = x.double().requires_grad_(True)
    yr = F.conv2d(xr, w.double(), bias.double() if bias is not None else None, stride=stride, padding=k // 2)
    gy = O.det_normal(tag + "gy", tuple(yr.shape))
    yr.backward(gy.double())
    y32 = F.conv2d(x, w, bias, stride=stride, padding=k // 2)
    Ho, Wo = yr.shape[-2:]
    xa, Cin_p = to_act(x, "fp32")
    Cout_p = pad(Cout, 4)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin_p, ldx=Cin_p, Ho=Ho, Wo=Wo, Cout_p=Cout_p, ldy=Cout_p, k=k, stride=stride, pad=k // 2)
    assert N.lib().vae2_conv2d_tf32_supported(C.byref(g)) == 1
    Nf, Kf, NfT, KfT = N.tf32_dims(g)
    wd = w.contiguous().to(dev())
    wf = torch.zeros(3 * k * k * Nf * Kf, dtype=torch.bfloat16, device=dev())
    wb = torch.zeros(3 * k * k * NfT * KfT, dtype=torch.bfloat16, device=dev())
    d = (N.Tf32PackDesc * 1)()
    d[0] = N.Tf32PackDesc(w=wd.data_ptr(), fwd=wf.data_ptr(), bwd=wb.data_ptr(), Cout=Cout, Cin=Cin, k=k, Nf=Nf, Kf=Kf, NfT=NfT, KfT=KfT)
    t = table(d)
    N.call.vae2_pack_weights_tf32(t.data_ptr(), 1, st())
    ya = torch.zeros(B * Ho * Wo * Cout_p, dtype=torch.float32, device=dev())
    bp = None
    if bias is not None:
        bp = torch.zeros(Cout_p, dtype=torch.float32, device=dev())
        bp[:Cout] = bias.to(dev())
    N.call.vae2_conv2d_fwd(xa.data_ptr(), wf.data_ptr(), bp.data_ptr() if bp is not None else None, ya.data_ptr(), 0,
                           C.byref(g), 2, st())
    torch.cuda.synchronize()
    y = from_act(ya, "fp32", B, Cout, Ho, Wo, Cout_p)
    e_mine, e_ref = rel_err(y, yr.detach()), rel_err(y32, yr.detach())
    from helpers import log_err
    log_err("f32x3_conv_%s" % (case,), f32x3_vs_fp64=e_mine, torch_fp32_vs_fp64=e_ref)
    assert e_mine < 3e-5, "f32x3 fwd err %.2e vs torch-fp32 err %.2e" % (e_mine, e_ref)
    assert float(ya.view(-1, Cout_p)[:, Cout:].abs().sum()) == 0.0
    gya, _ = to_act(gy, "fp32", Cout_p)
    dxa = torch.zeros_like(xa)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wb.data_ptr(), dxa.data_ptr(), 0, C.byref(g), 0, 2, st())
    torch.cuda.synchronize()
    dx = from_act(dxa, "fp32", B, Cin, H, W, Cin_p)
    assert rel_err(dx, xr.grad) < 3e-5, "f32x3 dgrad err %.2e" % rel_err(dx, xr.grad)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wb.data_ptr(), dxa.data_ptr(), 0, C.byref(g), 1, 2, st())
    torch.cuda.synchronize()
    assert rel_err(from_act(dxa, "fp32", B, Cin, H, W, Cin_p), 2 * xr.grad) < 3e-5, "f32x3 dgrad accumulate"


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_wgrad_f32x2(case):
    """The fp32 path's tensor-core weight gradient (bf16 hi/lo planes, three products on the tcgen05 wgrad kernel with
    separate accumulators) against a float64 weight gradient; the CUDA-core split-K kernel it replaces is measured
    beside it."""
    from helpers import log_err
    B, Cin, Cout, H, W, k, use_bias, stride = case
    tag = "wx2%s" % (case,)
    x = O.det_normal(tag + "x", (B, Cin, H, W))
    w = O.det_normal(tag + "w", (Cout, Cin, k, k), 0.05).double().requires_grad_(True)
    yr = F.conv2d(x.double(), w, None, stride=stride, padding=k // 2)
    gy = O.det_normal(tag + "gy", tuple(yr.shape))
    yr.backward(gy.double())
    Ho, Wo = yr.shape[-2:]
    xa, Cin_p = to_act(x, "fp32")
    gya, Cout_p = to_act(gy, "fp32")
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin_p, ldx=Cin_p, Ho=Ho, Wo=Wo, Cout_p=Cout_p, ldy=Cout_p, k=k, stride=stride, pad=k // 2)
    need = N.lib().vae2_conv2d_wgrad_f32x2_workspace(C.byref(g))
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device=dev())
    dwp = torch.full((k * k * Cin_p * Cout_p,), 7.0, dtype=torch.float32, device=dev())      # must be overwritten
    N.call.vae2_conv2d_wgrad_f32x2(xa.data_ptr(), gya.data_ptr(), dwp.data_ptr(), ws.data_ptr(), C.byref(g), st())
    dws = torch.zeros_like(dwp)
    N.call.vae2_conv2d_wgrad(xa.data_ptr(), gya.data_ptr(), dws.data_ptr(), 0, C.byref(g), 0, st())
    torch.cuda.synchronize()

    def unpack(t):       # [tap][Cin_p][Cout_p] -> OIHW
        return t.view(k, k, Cin_p, Cout_p)[:, :, :Cin, :Cout].permute(3, 2, 0, 1).cpu()
    e_tc, e_simt = rel_err(unpack(dwp), w.grad), rel_err(unpack(dws), w.grad)
    log_err("wgrad_f32x2_%s" % (case,), f32x2_vs_fp64=e_tc, cuda_core_vs_fp64=e_simt)
    assert e_tc < 3e-5, (e_tc, e_simt)
    pads = dwp.view(k * k, Cin_p, Cout_p)
    assert float(pads[:, Cin:, :].abs().sum()) == 0.0 and float(pads[:, :, Cout:].abs().sum()) == 0.0


@pytest.mark.parametrize("slot", [0, 1, 2])
def test_conv_tc_prediction_head_writes_an_8_lane_slice(slot):
    """The 270 -> 3 prediction heads (lib/models/enc_hrnet.py:324-337, last_layer) write an 8-lane slice of the 9-frame clip
    buffer: Cout_p = 8 inside ldy = 32.  On the tcgen05 path N is padded to 16 by TMA zero fill and the epilogues store 8
    lanes -- the neighbouring slices must stay untouched; forward (+bias), data gradient and weight gradient against torch."""
    B, Cin, Cout, H, W, k = 2, 270, 3, 19, 23, 1
    tag = "head%d" % slot
    x = O.det_normal(tag + "x", (B, Cin, H, W)).bfloat16().float()
    w = O.det_normal(tag + "w", (Cout, Cin, k, k), (2.0 / Cin) ** 0.5).bfloat16().float()
    bias = O.det_normal(tag + "b", (Cout,), 0.1)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, bias)
    gy = O.det_normal(tag + "gy", tuple(yr.shape)).bfloat16().float()
    yr.backward(gy)
    xa, Cin_p = to_act(x, "bf16", pad(Cin, 16))
    Cout_p, ldy = 8, 32
    wq, wqT = pack_bf16(w, Cin_p, Cout_p)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin_p, ldx=Cin_p, Ho=H, Wo=W, Cout_p=Cout_p, ldy=ldy, k=k, stride=1, pad=0)
    assert N.lib().vae2_conv2d_tc_supported(C.byref(g)) == 1
    clip = torch.full((B * H * W, ldy), 3.0, dtype=torch.bfloat16, device=dev())
    bp = torch.zeros(Cout_p, dtype=torch.float32, device=dev())
    bp[:Cout] = bias.to(dev())
    off = slot * 8
    N.call.vae2_conv2d_fwd(xa.data_ptr(), wq.data_ptr(), bp.data_ptr(), clip.data_ptr() + 2 * off, 1, C.byref(g), 1, st())
    torch.cuda.synchronize()
    y = clip[:, off:off + Cout].float().reshape(B, H, W, Cout).permute(0, 3, 1, 2).cpu()
    assert rel_err(y, yr.detach()) < 6e-3, "head fwd rel err %.3e" % rel_err(y, yr.detach())
    mask = torch.ones(ldy, dtype=torch.bool)
    mask[off:off + 8] = False
    assert bool((clip[:, mask.to(dev())] == 3.0).all()), "neighbouring slices were overwritten"
    assert float(clip[:, off + Cout:off + 8].float().abs().sum()) == 0.0
    # backward: dy lives in the same kind of slice, other lanes hold garbage that must not leak in
    gclip = torch.full((B * H * W, ldy), 5.0, dtype=torch.bfloat16, device=dev())
    gclip[:, off:off + 8] = 0
    gclip[:, off:off + Cout] = gy.permute(0, 2, 3, 1).reshape(-1, Cout).to(dev()).bfloat16()
    dxa = torch.zeros_like(xa)
    N.call.vae2_conv2d_dgrad(gclip.data_ptr() + 2 * off, wqT.data_ptr(), dxa.data_ptr(), 1, C.byref(g), 0, 1, st())
    torch.cuda.synchronize()
    dx = from_act(dxa, "bf16", B, Cin, H, W, Cin_p)
    assert rel_err(dx, xr.grad) < 6e-3, "head dgrad rel err %.3e" % rel_err(dx, xr.grad)
    need = N.lib().vae2_conv2d_wgrad_tc_workspace(C.byref(g))
    assert need > 0
    ws = torch.zeros(need, dtype=torch.float32, device=dev())
    dwp = torch.full((Cin_p * Cout_p + 64,), 7.0, dtype=torch.float32, device=dev())
    N.call.vae2_conv2d_wgrad_tc(xa.data_ptr(), gclip.data_ptr() + 2 * off, dwp.data_ptr(), ws.data_ptr(), C.byref(g), st())
    dw = torch.zeros_like(w).to(dev())
    d = (N.PackDesc * 1)()
    d[0] = N.PackDesc(w=dw.data_ptr(), wp=dwp.data_ptr(), Cout=Cout, Cin=Cin, k=k, Cin_p=Cin_p, Cout_p=Cout_p)
    t = table(d)
    N.call.vae2_unpack_wgrad(t.data_ptr(), 1, 0, st())
    torch.cuda.synchronize()
    assert rel_err(dw.cpu(), wr.grad) < 6e-3, "head wgrad rel err %.3e" % rel_err(dw.cpu(), wr.grad)
    assert bool((dwp[Cin_p * Cout_p:] == 7.0).all()), "weight-gradient rows overflowed"
