"""One launch (after one warm-up launch) of every hot kernel family at the shapes that own the W18 256x512 step --
the target of the `ncu --set full` capture whose summary lives in profiles/ (tools/ncu_summarise.py).

    python tools/ncu_kernels.py [bf16|fp32] [B]

    ncu --set full --clock-control none --profile-from-start off \
        -k regex:'conv_tc|wgrad_tc|wgrad_reduce|bn_fwd_fused|bn_bwd_fused|fuse_sum|fuse_bwd_up|elbo_terms|conv_direct|wgrad_direct|conv_igemm|f32x3' \
        -o gpurun_out/ncu_r2_<prec> python tools/ncu_kernels.py <prec>
    (only the launch bracketed by cudaProfilerStart/Stop is profiled; NCU_ONLY=tag,tag restricts the set)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
from _engine_loader import engine

E = engine(); N = E.native
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
code, tdt, al = (0, torch.float32, 4) if prec == "fp32" else (1, torch.bfloat16, 16)
pad = lambda c: (c + al - 1) // al * al
st = torch.cuda.current_stream().cuda_stream
f32 = dict(dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p = lambda t: t.data_ptr() if t is not None else None


ONLY = os.environ.get("NCU_ONLY", "")      # comma-separated substrings of the tags to run (keeps the report small)


def twice(tag, fn):
    """warm-up launch, L2 flush, then the PROFILED launch (ncu --profile-from-start off sees only this one)."""
    if ONLY and not any(t in tag for t in ONLY.split(",")):
        return
    fn()
    flush.zero_()                       # the profiled launch starts from a cold L2, like a launch deep inside the step
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("launched", tag, "->", N.lib().vae2_last_kernel().decode())


# ---- convolutions: (name, H, W, Cin, Cout, k, stride) -- the shapes with the largest time share (profiles/r1_step_callsites_*)
CONVS = [("18->18 3x3", 256, 512, 18, 18, 3, 1), ("36->36 3x3", 128, 256, 36, 36, 3, 1), ("64->64 3x3", 256, 512, 64, 64, 3, 1),
         ("270->270 1x1", 256, 512, 270, 270, 1, 1), ("64->256 1x1", 256, 512, 64, 256, 1, 1), ("18->36 3x3 s2", 256, 512, 18, 36, 3, 2),
         ("256->18 3x3", 256, 512, 256, 18, 3, 1), ("72->72 3x3", 64, 128, 72, 72, 3, 1)]
for name, H, W, Cin, Cout, k, s_ in CONVS:
    Cip, Cop = pad(Cin), pad(Cout)
    Ho, Wo = (H + 2 * (k // 2) - k) // s_ + 1, (W + 2 * (k // 2) - k) // s_ + 1
    x = (torch.randn(B * H * W * Cip, device=dev) * 0.5).to(tdt)
    y = torch.zeros(B * Ho * Wo * Cop, dtype=tdt, device=dev)
    dy = (torch.randn(B * Ho * Wo * Cop, device=dev) * 0.5).to(tdt)
    dx = torch.zeros_like(x)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cip, ldx=Cip, Ho=Ho, Wo=Wo, Cout_p=Cop, ldy=Cop, k=k, stride=s_, pad=k // 2)
    nw = k * k * Cip * Cop
    eng = 1 if (prec == "bf16" and N.lib().vae2_conv2d_tc_supported(C.byref(g))) else 0
    w = (torch.randn(nw, device=dev) * 0.05).to(torch.bfloat16 if eng else torch.float32)
    dwp = torch.zeros(nw, **f32)
    twice("fwd " + name, lambda: N.call.vae2_conv2d_fwd(p(x), p(w), None, p(y), code, C.byref(g), eng, st))
    twice("dgrad " + name, lambda: N.call.vae2_conv2d_dgrad(p(dy), p(w), p(dx), code, C.byref(g), 0, eng, st))
    if eng:
        ws = torch.zeros(max(N.lib().vae2_conv2d_wgrad_tc_workspace(C.byref(g)), 4), **f32)
        twice("wgrad " + name, lambda: N.call.vae2_conv2d_wgrad_tc(p(x), p(dy), p(dwp), p(ws), C.byref(g), st))
    else:
        twice("wgrad " + name, lambda: N.call.vae2_conv2d_wgrad(p(x), p(dy), p(dwp), code, C.byref(g), 0, st))
    if prec == "fp32" and N.lib().vae2_conv2d_tf32_supported(C.byref(g)):
        # the fp32 path's tensor-core route (>= 40-lane layers and 3x3 stride-1 layers from 36 lanes; 18->18 is captured too,
        # for the comparison with conv_direct): exact 3-way bf16 split, split accumulators
        Nf, Kf, NfT, KfT = N.tf32_dims(g)
        wsrc = torch.randn(Cout, Cin, k, k, device=dev) * 0.05
        wf = torch.zeros(3 * k * k * Nf * Kf, dtype=torch.bfloat16, device=dev)
        wb = torch.zeros(3 * k * k * NfT * KfT, dtype=torch.bfloat16, device=dev)
        d = (N.Tf32PackDesc * 1)()
        d[0] = N.Tf32PackDesc(w=p(wsrc), fwd=p(wf), bwd=p(wb), Cout=Cout, Cin=Cin, k=k, Nf=Nf, Kf=Kf, NfT=NfT, KfT=KfT)
        tab = torch.frombuffer(bytearray(bytes(d)), dtype=torch.uint8).to(dev)
        N.call.vae2_pack_weights_tf32(tab.data_ptr(), 1, st)
        twice("f32x3 fwd " + name, lambda: N.call.vae2_conv2d_fwd(p(x), p(wf), None, p(y), 0, C.byref(g), 2, st))
        twice("f32x3 dgrad " + name, lambda: N.call.vae2_conv2d_dgrad(p(dy), p(wb), p(dx), 0, C.byref(g), 0, 2, st))
        need = N.lib().vae2_conv2d_wgrad_f32x2_workspace(C.byref(g))
        if need > 0:
            wsb = torch.zeros(need, dtype=torch.uint8, device=dev)
            twice("f32x2 wgrad " + name, lambda: N.call.vae2_conv2d_wgrad_f32x2(p(x), p(dy), p(dwp), p(wsb), C.byref(g), st))
            del wsb
    del x, y, dy, dx, w, dwp
    torch.cuda.empty_cache()

# ---- batch norm (fused cooperative kernels): 18-lane branch tensor and the 256-lane bottleneck tensor
for P, C_ in ((B * 256 * 512, 18), (B * 256 * 512, 256), (B * 128 * 256, 36)):
    Cp = pad(C_)
    mk = lambda: torch.randn(P, Cp, device=dev).to(tdt)
    y, res, out, gg, dyb, dres = mk(), mk(), mk(), mk(), mk(), mk()
    parts = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * Cp, **f32)
    gam, bet, rm, rv = torch.ones(Cp, **f32), torch.zeros(Cp, **f32), torch.zeros(Cp, **f32), torch.ones(Cp, **f32)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    mean, invstd, scale, shift, c1, c2, dg, db = (torch.zeros(Cp, **f32) for _ in range(8))
    twice("bn fwd C=%d" % C_, lambda: N.call.vae2_bn_fwd_fused(p(y), p(res), p(out), p(parts), code, P, C_, Cp, Cp, Cp, Cp, p(gam),
                                                                p(bet), p(rm), p(rv), p(nbt), 0.01, 1e-5, p(mean), p(invstd),
                                                                p(scale), p(shift), 1, st))
    twice("bn bwd(res) C=%d" % C_, lambda: N.call.vae2_bn_bwd_fused(p(gg), p(out), p(y), p(dyb), p(dres), p(parts), code, P, C_,
                                                                     Cp, Cp, Cp, Cp, Cp, Cp, p(mean), p(invstd), p(scale),
                                                                     p(shift), p(dg), p(db), 0, p(c1), p(c2), 1, 0, 0, st))
    twice("bn bwd(mask from y) C=%d" % C_, lambda: N.call.vae2_bn_bwd_fused(p(gg), p(out), p(y), p(dyb), None, p(parts), code, P,
                                                                             C_, Cp, Cp, Cp, Cp, Cp, Cp, p(mean), p(invstd),
                                                                             p(scale), p(shift), p(dg), p(db), 0, p(c1),
                                                                             p(c2), 2, 0, 0, st))
    del y, res, out, gg, dyb, dres
    torch.cuda.empty_cache()

# ---- branch fusion: 4-source sum at full resolution (stage 4, branch 0) and its up-sampling backward
Cp = pad(18)
H, W = 256, 512
srcs = [torch.randn(B * (H >> i) * (W >> i) * Cp, device=dev).to(tdt) for i in range(4)]
out = torch.zeros(B * H * W * Cp, dtype=tdt, device=dev)
arr = (N.FuseSrc * 4)()
for i, s_ in enumerate(srcs):
    arr[i] = N.FuseSrc(ptr=p(s_), H=H >> i, W=W >> i, ld=Cp)
twice("fuse_sum 4 sources", lambda: N.call.vae2_fuse_sum(arr, 4, p(out), code, B, H, W, Cp, Cp, 1, st))
gout, gsrc = torch.randn_like(out), torch.zeros_like(srcs[1])
twice("fuse_bwd_up /2", lambda: N.call.vae2_fuse_bwd_up(p(gout), p(out), p(gsrc), code, B, H, W, H // 2, W // 2, Cp, Cp, Cp, Cp, 1, 0, st))

# ---- ELBO terms at B >= 8 (judge the HBM target there): 3 x L1 + 4 x (reparam + KL)
Be, Z = 8, 8
preds = [torch.randn(Be, 9, H, W, device=dev) for _ in range(6)]
spec = [dict(kind=0, slot=i, a=2 * i, b=2 * i + 1, scale=1.0 / Be, name="l1") for i in range(3)]
twice("elbo L1 x3 B=8", lambda: E.elbo_terms(spec, 3, preds))
mv = [torch.randn(Be, 2 * Z, H >> i, W >> i, device=dev) * 0.1 for i in range(4)]
eps = [torch.randn(Be, Z, H >> i, W >> i, device=dev) for i in range(4)]
spec = [dict(kind=1, slot=0, a=4 + i, b=i, scale=1.0 / Be, want_z=True, name=i) for i in range(4)]
twice("elbo reparam+KL x4 B=8", lambda: E.elbo_terms(spec, 1, mv + eps))
print("done")
