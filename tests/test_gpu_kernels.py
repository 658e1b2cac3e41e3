"""GPU: every kernel family called through the C ABI (ctypes) against torch CPU references."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import E, O, rel_err

N = E.native
pytestmark = pytest.mark.gpu

DT = {"fp32": (0, torch.float32, 4, 2e-5), "bf16": (1, torch.bfloat16, 8, 2e-2)}


def dev():
    return torch.device("cuda:0")


def st():
    return torch.cuda.current_stream().cuda_stream


def pad(c, a):
    return (c + a - 1) // a * a


def to_act(x, prec, Cp=None):
    """NCHW fp32 (cpu) -> channels-last padded device tensor via the ABI."""
    code, tdt, al, _ = DT[prec]
    B, C_, H, W = x.shape
    Cp = Cp or pad(C_, al)
    xd = x.contiguous().to(dev())
    out = torch.zeros(B * H * W * Cp, dtype=tdt, device=dev())
    N.call.vae2_nchw_to_act(xd.data_ptr(), out.data_ptr(), code, B, C_, Cp, H, W, Cp, C_, 0, st())
    return out, Cp


def from_act(a, prec, B, C_, H, W, Cp):
    code = DT[prec][0]
    out = torch.zeros(B, C_, H, W, dtype=torch.float32, device=dev())
    N.call.vae2_act_to_nchw(a.data_ptr(), out.data_ptr(), code, B, C_, H, W, Cp, C_, 0, 0, st())
    return out.cpu()


def table(structs):
    return torch.frombuffer(bytearray(bytes(structs)), dtype=torch.uint8).to(dev())


def pack(w, Cin_p, Cout_p, cin_map=None):
    Cout, Cin, k, _ = w.shape
    wd = w.contiguous().to(dev())
    n = k * k * Cin_p * Cout_p
    wp = torch.zeros(n, dtype=torch.float32, device=dev())
    wpT = torch.zeros(n, dtype=torch.float32, device=dev())
    cm = None
    if cin_map is not None:
        cm = torch.tensor(cin_map, dtype=torch.int32, device=dev())
    d = (N.PackDesc * 1)()
    d[0] = N.PackDesc(w=wd.data_ptr(), wp=wp.data_ptr(), wpT=wpT.data_ptr(), cin_map=cm.data_ptr() if cm is not None else None,
                      Cout=Cout, Cin=Cin, k=k, Cin_p=Cin_p, Cout_p=Cout_p)
    t = table(d)
    N.call.vae2_pack_weights(t.data_ptr(), 1, st())
    torch.cuda.synchronize()
    return wp, wpT


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_layout_roundtrip(prec):
    x = O.det_normal("layout", (2, 9, 7, 13))
    if prec == "bf16":
        x = x.bfloat16().float()
    a, Cp = to_act(x, prec)
    y = from_act(a, prec, 2, 9, 7, 13, Cp)
    assert torch.equal(x, y)
    v = a.view(2, 7, 13, Cp).float().cpu()
    assert torch.equal(v[..., :9].permute(0, 3, 1, 2), x)
    assert float(v[..., 9:].abs().sum()) == 0.0


CONV_CASES = [  # (B, Cin, Cout, H, W, k, s, bias)
    (2, 9, 64, 12, 20, 3, 1, False),
    (1, 18, 18, 17, 23, 3, 1, False),
    (2, 18, 36, 17, 23, 3, 2, False),
    (1, 64, 256, 9, 11, 1, 1, False),
    (2, 270, 3, 8, 8, 1, 1, True),
    (1, 36, 36, 33, 47, 3, 2, False),
    (1, 72, 144, 16, 16, 3, 2, False),
    (3, 144, 18, 5, 6, 1, 1, False),
    (2, 36, 18, 9, 40, 1, 1, False),      # 1x1 between narrow branches (fp32: direct fwd/dgrad/wgrad kernels)
    (1, 18, 72, 21, 26, 3, 2, False),     # stride 2, narrow -> 72 lanes, odd height (parity-class dgrad)
    (2, 36, 36, 40, 70, 3, 1, False),     # several patches per image, W not a multiple of the 32-column warp
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, prec):
    B, Cin, Cout, H, W, k, s, use_bias = case
    code, tdt, al, tol = DT[prec]
    tag = "conv%s" % (case,)
    x = O.det_normal(tag + "x", (B, Cin, H, W))
    w = O.det_normal(tag + "w", (Cout, Cin, k, k), (2.0 / (Cin * k * k)) ** 0.5)
    bias = O.det_normal(tag + "b", (Cout,), 0.1) if use_bias else None
    if prec == "bf16":
        x = x.bfloat16().float()
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, bias, stride=s, padding=k // 2)
    gy = O.det_normal(tag + "gy", tuple(yr.shape))
    if prec == "bf16":
        gy = gy.bfloat16().float()
    yr.backward(gy)

    xa, Cin_p = to_act(x, prec)
    Cout_p = pad(Cout, al)
    wp, wpT = pack(w, Cin_p, Cout_p)
    ya = torch.zeros(B * Ho * Wo * Cout_p, dtype=tdt, device=dev())
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin_p, ldx=Cin_p, Ho=Ho, Wo=Wo, Cout_p=Cout_p, ldy=Cout_p, k=k, stride=s, pad=k // 2)
    bp = None
    if bias is not None:
        bp = torch.zeros(Cout_p, dtype=torch.float32, device=dev())
        bp[:Cout] = bias.to(dev())
    N.call.vae2_conv2d_fwd(xa.data_ptr(), wp.data_ptr(), bp.data_ptr() if bp is not None else None, ya.data_ptr(), code,
                           C.byref(g), 0, st())
    y = from_act(ya, prec, B, Cout, Ho, Wo, Cout_p)
    assert rel_err(y, yr.detach()) < tol, "fwd"
    assert float(ya.view(-1, Cout_p)[:, Cout:].float().abs().sum()) == 0.0, "pad lanes must stay zero"

    gya, _ = to_act(gy, prec, Cout_p)
    dxa = torch.zeros_like(xa)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wpT.data_ptr(), dxa.data_ptr(), code, C.byref(g), 0, 0, st())
    dx = from_act(dxa, prec, B, Cin, H, W, Cin_p)
    assert rel_err(dx, xr.grad) < tol, "dgrad"
    # accumulate form: dx += conv_T(dy)
    N.call.vae2_conv2d_dgrad(gya.data_ptr(), wpT.data_ptr(), dxa.data_ptr(), code, C.byref(g), 1, 0, st())
    assert rel_err(from_act(dxa, prec, B, Cin, H, W, Cin_p), 2 * xr.grad) < tol, "dgrad accumulate"

    dwp = torch.zeros_like(wp)
    N.call.vae2_conv2d_wgrad(xa.data_ptr(), gya.data_ptr(), dwp.data_ptr(), code, C.byref(g), 0, st())
    dw = torch.zeros_like(w).to(dev())
    d = (N.PackDesc * 1)()
    d[0] = N.PackDesc(w=dw.data_ptr(), wp=dwp.data_ptr(), Cout=Cout, Cin=Cin, k=k, Cin_p=Cin_p, Cout_p=Cout_p)
    t = table(d)
    N.call.vae2_unpack_wgrad(t.data_ptr(), 1, 0, st())
    assert rel_err(dw.cpu(), wr.grad) < tol, "wgrad"
    if bias is not None:
        db = torch.zeros(Cout, dtype=torch.float32, device=dev())
        N.call.vae2_bias_grad(gya.data_ptr(), db.data_ptr(), code, B * Ho * Wo, Cout, Cout_p, 0, st())
        assert rel_err(db.cpu(), gy.sum((0, 2, 3))) < tol, "bias grad"


def test_conv_concat_channel_map():
    """A conv over a [8 | 8 | 18]-segment concat buffer with the logical->physical lane map."""
    B, H, W, k = 1, 9, 10, 3
    segs = [8, 8, 18]
    x = O.det_normal("catx", (B, sum(segs), H, W))
    w = O.det_normal("catw", (18, sum(segs), k, k), 0.1)
    yr = F.conv2d(x, w, None, 1, 1)
    al = 4
    seg_p = [pad(c, al) for c in segs]
    Cp = pad(sum(seg_p), al)
    buf = torch.zeros(B * H * W * Cp, dtype=torch.float32, device=dev())
    xd = x.to(dev())
    cmap, off, src_off = [], 0, 0
    for c, cp in zip(segs, seg_p):
        N.call.vae2_nchw_to_act(xd.data_ptr(), buf.data_ptr() + 4 * off, 0, B, c, cp, H, W, Cp, sum(segs), src_off, st())
        cmap += list(range(off, off + c))
        off += cp
        src_off += c
    Cout_p = pad(18, al)
    wp, _ = pack(w, Cp, Cout_p, cmap)
    ya = torch.zeros(B * H * W * Cout_p, dtype=torch.float32, device=dev())
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cp, ldx=Cp, Ho=H, Wo=W, Cout_p=Cout_p, ldy=Cout_p, k=k, stride=1, pad=1)
    N.call.vae2_conv2d_fwd(buf.data_ptr(), wp.data_ptr(), None, ya.data_ptr(), 0, C.byref(g), 0, st())
    assert rel_err(from_act(ya, "fp32", B, 18, H, W, Cout_p), yr) < 2e-5


@pytest.mark.parametrize("fused", [False, True], ids=["split", "fused"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape,relu,with_res", [((2, 18, 17, 23), True, True), ((1, 64, 32, 64), True, False),
                                                   ((3, 270, 5, 7), False, False), ((2, 8, 64, 64), True, True),
                                                   ((4, 36, 96, 160), True, True), ((1, 20, 3, 5), False, False)])
def test_batchnorm_train_fwd_bwd(shape, relu, with_res, prec, fused):
    code, tdt, al, tol = DT[prec]
    B, C_, H, W = shape
    tag = "bn%s" % (shape,)
    y = O.det_normal(tag + "y", shape, 2.0, 0.5)
    res = O.det_normal(tag + "r", shape) if with_res else None
    gam, bet = O.det_uniform(tag + "g", (C_,), 0.5, 1.5), O.det_normal(tag + "b", (C_,), 0.1)
    rm, rv = O.det_normal(tag + "rm", (C_,), 0.1), O.det_uniform(tag + "rv", (C_,), 0.5, 1.5)
    if prec == "bf16":
        y = y.bfloat16().float()
        res = res.bfloat16().float() if res is not None else None
    yr = y.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if res is not None else None
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    o = F.batch_norm(yr, rm_r, rv_r, gr, br, True, 0.01, 1e-5)
    if rr is not None:
        o = o + rr
    if relu:
        o = F.relu(o)
    go = O.det_normal(tag + "go", shape)
    if prec == "bf16":
        go = go.bfloat16().float()
    o.backward(go)

    P = B * H * W
    ya, Cp = to_act(y, prec)
    ra = to_act(res, prec)[0] if res is not None else None
    f32 = dict(dtype=torch.float32, device=dev())
    parts = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * Cp, **f32)
    npart = C.c_int(0)
    gd, bd, rmd, rvd = gam.to(dev()), bet.to(dev()), rm.to(dev()), rv.to(dev())
    nbt = torch.zeros(1, dtype=torch.int64, device=dev())
    mean, invstd, scale, shift = (torch.zeros(Cp, **f32) for _ in range(4))
    oa = torch.zeros_like(ya)
    if fused:
        N.call.vae2_bn_fwd_fused(ya.data_ptr(), ra.data_ptr() if ra is not None else None, oa.data_ptr(), parts.data_ptr(),
                                 code, P, C_, Cp, Cp, Cp, Cp, gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(), rvd.data_ptr(),
                                 nbt.data_ptr(), 0.01, 1e-5, mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(),
                                 shift.data_ptr(), 1 if relu else 0, st())
    else:
        N.call.vae2_bn_stats(ya.data_ptr(), parts.data_ptr(), C.byref(npart), code, P, Cp, Cp, st())
        N.call.vae2_bn_finalize(parts.data_ptr(), npart.value, C_, Cp, gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(),
                                rvd.data_ptr(), nbt.data_ptr(), 0.01, 1e-5, mean.data_ptr(), invstd.data_ptr(),
                                scale.data_ptr(), shift.data_ptr(), st())
        N.call.vae2_bn_apply(ya.data_ptr(), ra.data_ptr() if ra is not None else None, oa.data_ptr(), code, P, Cp, Cp, Cp,
                             Cp, scale.data_ptr(), shift.data_ptr(), 1 if relu else 0, st())
    assert rel_err(from_act(oa, prec, B, C_, H, W, Cp), o.detach()) < tol, "apply"
    assert rel_err(rmd.cpu(), rm_r) < 1e-5 and rel_err(rvd.cpu(), rv_r) < 1e-5, "running stats"
    assert int(nbt) == 1
    assert rel_err(mean[:C_].cpu(), y.mean((0, 2, 3))) < 1e-5

    ga, _ = to_act(go, prec)
    parts2 = torch.zeros(N.lib().vae2_bn_max_partials() * 2 * Cp, **f32)
    sums, c1, c2 = torch.zeros(2 * Cp, **f32), torch.zeros(Cp, **f32), torch.zeros(Cp, **f32)
    dg, db = torch.zeros(C_, **f32), torch.zeros(C_, **f32)
    # the dx / dres buffers start non-zero and are accumulated into: the plan's "second writer" case
    dy0 = O.det_normal(tag + "dy0", shape)
    dya, _ = to_act(dy0, prec)
    dy0 = from_act(dya, prec, B, C_, H, W, Cp)
    dra = dya.clone() if ra is not None else None
    if fused:
        N.call.vae2_bn_bwd_fused(ga.data_ptr(), oa.data_ptr(), ya.data_ptr(), dya.data_ptr(),
                                 dra.data_ptr() if dra is not None else None, parts2.data_ptr(), code, P, C_, Cp, Cp, Cp,
                                 Cp, Cp, Cp, mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                 dg.data_ptr(), db.data_ptr(), 0, c1.data_ptr(), c2.data_ptr(),
                                 0 if not relu else (1 if with_res else 2), 1, 1, st())
    else:
        N.call.vae2_bn_bwd_reduce(ga.data_ptr(), oa.data_ptr(), ya.data_ptr(), parts2.data_ptr(), C.byref(npart), code, P,
                                  Cp, Cp, Cp, Cp, mean.data_ptr(), invstd.data_ptr(), 1 if relu else 0, st())
        N.call.vae2_bn_bwd_finalize(parts2.data_ptr(), npart.value, C_, Cp, sums.data_ptr(), st())
        N.call.vae2_bn_bwd_coeffs(sums.data_ptr(), C_, Cp, 1.0 / P, dg.data_ptr(), db.data_ptr(), 0, sums.data_ptr(),
                                  c1.data_ptr(), c2.data_ptr(), st())
        N.call.vae2_bn_bwd_elemt(ga.data_ptr(), oa.data_ptr(), ya.data_ptr(), dya.data_ptr(),
                                 dra.data_ptr() if dra is not None else None, code, P, Cp, Cp, Cp, Cp, Cp, Cp,
                                 mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), c1.data_ptr(), c2.data_ptr(),
                                 1 if relu else 0, 1, 1, st())
    btol = tol * 5
    assert rel_err(from_act(dya, prec, B, C_, H, W, Cp) - dy0, yr.grad) < btol, "dx"
    assert rel_err(dg.cpu(), gr.grad) < btol and rel_err(db.cpu(), br.grad) < btol, "dgamma/dbeta"
    if dra is not None:
        assert rel_err(from_act(dra, prec, B, C_, H, W, Cp) - dy0, rr.grad) < btol, "dres"


def test_batchnorm_eval_coeffs():
    C_, Cp = 18, 20
    gam, bet = O.det_uniform("e:g", (C_,), 0.5, 1.5), O.det_normal("e:b", (C_,), 0.1)
    rm, rv = O.det_normal("e:rm", (C_,), 0.1), O.det_uniform("e:rv", (C_,), 0.5, 1.5)
    f32 = dict(dtype=torch.float32, device=dev())
    sc, sh = torch.zeros(Cp, **f32), torch.zeros(Cp, **f32)
    t = [v.to(dev()) for v in (gam, bet, rm, rv)]
    N.call.vae2_bn_eval_coeffs(C_, Cp, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), 1e-5,
                               sc.data_ptr(), sh.data_ptr(), st())
    ref_sc = gam / torch.sqrt(rv + 1e-5)
    assert rel_err(sc[:C_].cpu(), ref_sc) < 1e-6 and rel_err(sh[:C_].cpu(), bet - rm * ref_sc) < 1e-6
    assert float(sc[C_:].abs().sum()) == 0.0


FUSE_CASES = [((32, 64), [(32, 64), (16, 32), (8, 16), (4, 8)]), ((33, 47), [(33, 47), (17, 24), (9, 12)]),
              ((17, 24), [(17, 24), (17, 24), (9, 12)]), ((60, 60), [(60, 60), (119, 119)][:1] + [(30, 30)])]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", FUSE_CASES)
def test_fuse_sum_fwd_bwd(case, prec):
    code, tdt, al, tol = DT[prec]
    (H, W), sizes = case
    B, C_ = 2, 18
    srcs = [O.det_normal("fuse%s%d" % (case, i), (B, C_, h, w)) for i, (h, w) in enumerate(sizes)]
    if prec == "bf16":
        srcs = [s.bfloat16().float() for s in srcs]
    sr = [s.clone().requires_grad_(True) for s in srcs]
    tot = None
    for s in sr:
        t = s if tuple(s.shape[-2:]) == (H, W) else F.interpolate(s, size=(H, W), mode="bilinear", align_corners=False)
        tot = t if tot is None else tot + t
    o = F.relu(tot)
    go = O.det_normal("fusego%s" % (case,), tuple(o.shape))
    if prec == "bf16":
        go = go.bfloat16().float()
    o.backward(go)
    acts = [to_act(s, prec) for s in srcs]
    Cp = acts[0][1]
    arr = (N.FuseSrc * len(srcs))()
    for i, ((a, _), (h, w)) in enumerate(zip(acts, sizes)):
        arr[i] = N.FuseSrc(ptr=a.data_ptr(), H=h, W=w, ld=Cp)
    oa = torch.zeros(B * H * W * Cp, dtype=tdt, device=dev())
    N.call.vae2_fuse_sum(arr, len(srcs), oa.data_ptr(), code, B, H, W, Cp, Cp, 1, st())
    assert rel_err(from_act(oa, prec, B, C_, H, W, Cp), o.detach()) < tol, "fwd"
    ga, _ = to_act(go, prec)
    for i, ((a, _), (h, w)) in enumerate(zip(acts, sizes)):
        gs = torch.zeros_like(a)
        if (h, w) == (H, W):
            d = (N.FuseDst * 1)()
            d[0] = N.FuseDst(ptr=gs.data_ptr(), ld=Cp, accumulate=0)
            N.call.vae2_fuse_bwd_same(ga.data_ptr(), oa.data_ptr(), d, 1, code, B * H * W, Cp, Cp, Cp, 1, st())
        else:
            N.call.vae2_fuse_bwd_up(ga.data_ptr(), oa.data_ptr(), gs.data_ptr(), code, B, H, W, h, w, Cp, Cp, Cp, Cp, 1, 0, st())
        assert rel_err(from_act(gs, prec, B, C_, h, w, Cp), sr[i].grad) < tol * 2, "bwd src %d" % i


def test_elbo_terms_kernel_matches_oracle():
    B, Z = 2, 8
    sizes = O.branch_sizes(17, 23)
    muvars = [O.det_normal("el:mv%d" % i, (B, 2 * Z, h, w), 0.3) for i, (h, w) in enumerate(sizes)]
    eps = [O.det_normal("el:eps%d" % i, (B, Z, h, w)) for i, (h, w) in enumerate(sizes)]
    preds = [O.det_normal("el:p%d" % i, (B, 9, 17, 23)) for i in range(3)]
    tgts = [O.det_normal("el:t%d" % i, (B, 9, 17, 23)) for i in range(3)]
    samp = O.det_normal("el:s", (B, 1, 17, 23))
    mv_r = [m.clone().requires_grad_(True) for m in muvars]
    pr_r = [p.clone().requires_grad_(True) for p in preds]
    sm_r = samp.clone().requires_grad_(True)
    mus, lvs = [m[:, :Z] for m in mv_r], [m[:, Z:] for m in mv_r]
    z_r = O.reparam(mus, lvs, eps)
    kl_r = O.kl_loss(mus, lvs)
    l1_r = [O.l1_loss(p, t) for p, t in zip(pr_r, tgts)]
    gan_r = 0.5 * O.lsgan_loss(sm_r, "real")
    zw = [O.det_normal("el:zw%d" % i, tuple(z.shape)) for i, z in enumerate(z_r)]
    total_r = 1.3 * kl_r + sum(float(i + 1) * l for i, l in enumerate(l1_r)) + 0.7 * gan_r + sum((z * w).sum() for z, w in zip(z_r, zw))
    total_r.backward()

    d = dev()
    mv = [m.to(d).requires_grad_(True) for m in muvars]
    ep = [e.to(d) for e in eps]
    n = len(mv)
    spec = [dict(kind=1, slot=0, a=n + i, b=i, scale=1.0 / B, want_z=True, name=i) for i in range(n)]
    out = E.elbo_terms(spec, 1, mv + ep)
    kl, z = out[0][0], out[1:]
    assert rel_err(kl, kl_r.detach()) < 1e-5
    for a, b in zip(z, z_r):
        assert rel_err(a, b.detach()) < 1e-6
    pr = [p.to(d).requires_grad_(True) for p in preds]
    tg = [t.to(d) for t in tgts]
    sm = samp.to(d).requires_grad_(True)
    spec = [dict(kind=0, slot=i, a=2 * i, b=2 * i + 1, scale=1.0 / B) for i in range(3)]
    spec.append(dict(kind=2, slot=3, a=6, b=None, scale=0.5 / B, target=1.0))
    vals = E.elbo_terms(spec, 4, [pr[0], tg[0], pr[1], tg[1], pr[2], tg[2], sm])[0]
    for i in range(3):
        assert rel_err(vals[i], l1_r[i].detach()) < 1e-5
    assert rel_err(vals[3], gan_r.detach()) < 1e-5
    total = 1.3 * kl + sum(float(i + 1) * vals[i] for i in range(3)) + 0.7 * vals[3] + sum((a * w.to(d)).sum() for a, w in zip(z, zw))
    total.backward()
    for a, b in zip(mv, mv_r):
        assert rel_err(a.grad, b.grad) < 1e-5, "d muvar"
    for a, b in zip(pr, pr_r):
        assert rel_err(a.grad, b.grad) < 1e-6, "d predict"
    assert rel_err(sm.grad, sm_r.grad) < 1e-5
    E.check_finite(block=True)


def test_elbo_nonfinite_is_reported():
    d = dev()
    p = torch.zeros(1, 9, 4, 4, device=d)
    p[0, 0, 0, 0] = float("nan")
    E.elbo_terms([dict(kind=0, slot=0, a=0, b=1, scale=1.0, name="x2t_predict")], 1, [p, torch.zeros_like(p)])
    with pytest.raises(AssertionError, match="x2t_predict got nan or inf"):
        E.check_finite(block=True)


def test_adam_step_matches_torch():
    n = 1003
    p0, g = O.det_normal("adam:p", (n,)), O.det_normal("adam:g", (n,))
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    d = dev()
    p, m, v = p0.to(d), torch.zeros(n, device=d), torch.zeros(n, device=d)
    step = torch.zeros(1, dtype=torch.int64, device=d)
    gd = g.to(d)
    for it in range(3):
        pr.grad = g.clone() * (it + 1)
        opt.step()
        step += 1
        N.call.vae2_adam_step(p.data_ptr(), (gd * (it + 1)).data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999,
                              1e-8, 0.0, step.data_ptr(), 1.0, st())
    assert rel_err(p.cpu(), pr.detach()) < 1e-6
