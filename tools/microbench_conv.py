"""Conv kernels timed alone (CUDA events on the launch stream, L2 flushed between launches).
Usage: python tools/microbench_conv.py [bf16|fp32] [reps]   -- also the target of the ncu --set full capture."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
from _engine_loader import engine

E = engine(); N = E.native
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
code, tdt, al = (0, torch.float32, 4) if prec == "fp32" else (1, torch.bfloat16, 16)
pad = lambda c: (c + al - 1) // al * al
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream(dev)
PEAK = 1662.8

SHAPES = [  # (name, B, H, W, Cin, Cout, k, stride)
    ("64->64 3x3 @256x512 (stem/bottleneck, 16% of W18 MACs)", 1, 256, 512, 64, 64, 3, 1),
    ("270->270 1x1 @256x512 (heads, 24%)", 1, 256, 512, 270, 270, 1, 1),
    ("18->18 3x3 @256x512 (branch 0, 10%)", 1, 256, 512, 18, 18, 3, 1),
    ("36->36 3x3 @128x256 (branch 1, 10%)", 1, 128, 256, 36, 36, 3, 1),
    ("64->256 1x1 @256x512 (bottleneck expand, 7%)", 1, 256, 512, 64, 256, 1, 1),
    ("144->144 3x3 @32x64 (branch 3)", 1, 32, 64, 144, 144, 3, 1),
    ("18->36 3x3 s2 @256x512 (fuse-down)", 1, 256, 512, 18, 36, 3, 2),
]


def timed(fn):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
        for a, b in ev:
            if not os.environ.get('VAE2_NOFLUSH'):
                flush.zero_()
            a.record(s); fn(); b.record(s)
    s.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]


for name, B, H, W, Cin, Cout, k, st in SHAPES:
    Cip, Cop = pad(Cin), pad(Cout)
    Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
    x = (torch.randn(B * H * W * Cip, device=dev) * 0.5).to(tdt)
    y = torch.zeros(B * Ho * Wo * Cop, dtype=tdt, device=dev)
    dy = (torch.randn(B * Ho * Wo * Cop, device=dev) * 0.5).to(tdt)
    dx = torch.zeros_like(x)
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cip, ldx=Cip, Ho=Ho, Wo=Wo, Cout_p=Cop, ldy=Cop, k=k, stride=st, pad=k // 2)
    nw = k * k * Cip * Cop
    flop = 2.0 * B * Ho * Wo * Cin * Cout * k * k           # algorithmic: no credit for lane padding
    eng = 1 if (prec == "bf16" and N.lib().vae2_conv2d_tc_supported(C.byref(g))) else 0
    w = (torch.randn(nw, device=dev) * 0.05).to(torch.bfloat16 if eng else torch.float32)
    sp = s.cuda_stream
    t_f = timed(lambda: N.call.vae2_conv2d_fwd(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), code, C.byref(g), eng, sp))
    t_d = timed(lambda: N.call.vae2_conv2d_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), code, C.byref(g), 0, eng, sp))
    dwp = torch.zeros(nw, dtype=torch.float32, device=dev)
    if eng:
        need = N.lib().vae2_conv2d_wgrad_tc_workspace(C.byref(g))
        ws = torch.zeros(max(need, 4), dtype=torch.float32, device=dev)
        t_w = timed(lambda: N.call.vae2_conv2d_wgrad_tc(x.data_ptr(), dy.data_ptr(), dwp.data_ptr(), ws.data_ptr(), C.byref(g), sp))
    else:
        t_w = timed(lambda: N.call.vae2_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dwp.data_ptr(), code, C.byref(g), 0, sp))
    tf = lambda ms: flop / (ms * 1e-3) / 1e12
    print("%-58s %s  fwd %7.1f us %6.1f TF (%4.1f%%) | dgrad %7.1f us %6.1f TF | wgrad %7.1f us %6.1f TF" % (
        name, "tc  " if eng else "simt", t_f * 1e3, tf(t_f), 100 * tf(t_f) / PEAK, t_d * 1e3, tf(t_d), t_w * 1e3, tf(t_w)))
