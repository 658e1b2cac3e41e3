"""BatchNorm kernels against the HBM roofline: split (stats | finalize | apply, reduce | finalize | coeffs | elemt)
and fused (one cooperative launch per direction) on the activation shapes of the W18 256x512 step.
Buffers rotate through > 2x the L2 capacity, so the figures are HBM figures, not L2 ones."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
from _engine_loader import engine

E = engine()
N = E.native
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
code, tdt, esz = (0, torch.float32, 4) if prec == "fp32" else (1, torch.bfloat16, 2)
dev = torch.device("cuda:0")
peak = 6531.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
SHAPES = [(524288, 256), (524288, 32), (131072, 48), (32768, 80), (8192, 144)]
if os.environ.get("BN_SHAPES"):
    SHAPES = [tuple(int(v) for v in t.split("x")) for t in os.environ["BN_SHAPES"].split(",")]
st = torch.cuda.current_stream().cuda_stream
f32 = dict(dtype=torch.float32, device=dev)


def timeit(fn, nbuf, iters=20):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


print("precision %s; HBM peak %.0f GB/s (MEASURED_PEAKS.json)" % (prec, peak))
for P, Cp in SHAPES:
    nbytes = P * Cp * esz
    nbuf = max(2, int(300e6 // (4 * nbytes)) + 1)
    mk = lambda: [torch.randn(P, Cp, device=dev).to(tdt) for _ in range(nbuf)]
    y, res, out, g, dy, dres = mk(), mk(), mk(), mk(), mk(), mk()
    parts = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * Cp, **f32)
    gam, bet, rm, rv = torch.ones(Cp, **f32), torch.zeros(Cp, **f32), torch.zeros(Cp, **f32), torch.ones(Cp, **f32)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    mean, invstd, scale, shift, c1, c2 = (torch.zeros(Cp, **f32) for _ in range(6))
    sums, dg, db = torch.zeros(2 * Cp, **f32), torch.zeros(Cp, **f32), torch.zeros(Cp, **f32)
    npart = C.c_int(0)
    p = lambda t: t.data_ptr()

    def split_fwd(i):
        N.call.vae2_bn_stats(p(y[i]), p(parts), C.byref(npart), code, P, Cp, Cp, st)
        N.call.vae2_bn_finalize(p(parts), npart.value, Cp, Cp, p(gam), p(bet), p(rm), p(rv), p(nbt), 0.01, 1e-5, p(mean),
                                p(invstd), p(scale), p(shift), st)
        N.call.vae2_bn_apply(p(y[i]), p(res[i]), p(out[i]), code, P, Cp, Cp, Cp, Cp, p(scale), p(shift), 1, st)

    def fused_fwd(i):
        N.call.vae2_bn_fwd_fused(p(y[i]), p(res[i]), p(out[i]), p(parts), code, P, Cp, Cp, Cp, Cp, Cp, p(gam), p(bet), p(rm),
                                 p(rv), p(nbt), 0.01, 1e-5, p(mean), p(invstd), p(scale), p(shift), 1, st)

    def split_bwd(i):
        N.call.vae2_bn_bwd_reduce(p(g[i]), p(out[i]), p(y[i]), p(parts), C.byref(npart), code, P, Cp, Cp, Cp, Cp, p(mean),
                                  p(invstd), 1, st)
        N.call.vae2_bn_bwd_finalize(p(parts), npart.value, Cp, Cp, p(sums), st)
        N.call.vae2_bn_bwd_coeffs(p(sums), Cp, Cp, 1.0 / P, p(dg), p(db), 0, p(sums), p(c1), p(c2), st)
        N.call.vae2_bn_bwd_elemt(p(g[i]), p(out[i]), p(y[i]), p(dy[i]), p(dres[i]), code, P, Cp, Cp, Cp, Cp, Cp, Cp, p(mean),
                                 p(invstd), p(scale), p(c1), p(c2), 1, 0, 0, st)

    def fused_bwd(i, mode=1):
        N.call.vae2_bn_bwd_fused(p(g[i]), p(out[i]), p(y[i]), p(dy[i]), p(dres[i]) if mode == 1 else None, p(parts), code, P,
                                 Cp, Cp, Cp, Cp, Cp, Cp, Cp, p(mean), p(invstd), p(scale), p(shift), p(dg), p(db), 0, p(c1),
                                 p(c2), mode, 0, 0, st)

    def fused_bwd_nores(i):
        fused_bwd(i, 2)

    # algorithmic bytes: fwd reads y (twice in the split path's two kernels, once + L2/again in the fused) + res, writes out
    fwd_bytes, bwd_bytes = 4 * nbytes, 8 * nbytes   # y,y,res,out | g,a,y (x2 passes), dy, dres
    for name, fn, nb in (("fwd split", split_fwd, fwd_bytes), ("fwd fused", fused_fwd, fwd_bytes),
                         ("bwd split", split_bwd, bwd_bytes), ("bwd fused", fused_bwd, bwd_bytes),
                         ("bwd fused, no residual (g,y | g,y,dy)", fused_bwd_nores, 5 * nbytes)):
        us = timeit(fn, nbuf)
        print("P=%7d Cp=%3d (%6.1f MB/tensor) %s %8.1f us  %7.0f GB/s  %5.1f%% of HBM peak" %
              (P, Cp, nbytes / 1e6, name, us, nb / us / 1e3, 100 * nb / us / 1e3 / peak))
    # phase split of the fused kernels (CTA 0's view)
    prof = torch.zeros(6, dtype=torch.int64, device=dev)
    N.call.vae2_debug_bn_phase_times(prof.data_ptr())
    for name, fn in (("fwd fused", fused_fwd), ("bwd fused", fused_bwd), ("bwd fused no-res", fused_bwd_nores)):
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        t = prof.cpu().tolist()
        print("    %-18s phases (us): pass1 %.1f | barrier %.1f | finalize %.1f | barrier %.1f | pass2 %.1f | total %.1f" %
              ((name,) + tuple((t[i + 1] - t[i]) / 1e3 for i in range(5)) + ((t[5] - t[0]) / 1e3,)))
    N.call.vae2_debug_bn_phase_times(None)
    del y, res, out, g, dy, dres
    torch.cuda.empty_cache()
