"""In-step kernel profile: every C-ABI launch of a training iteration bracketed by CUDA events on the launching stream,
grouped by (kernel, shape), each group carrying its ALGORITHMIC FLOPs and bytes -- so that bench.py can
name the kernel with the largest time share of the step and report its roofline fractions, instead of a hand-picked
best shape.

The step runs eagerly (no CUDA graph) while profiled; a device-side spin is queued ahead of every block of launches
so that the host stays ahead of the GPU and an event pair brackets kernel time, not Python launch latency.  The
numbers are in-situ ones: L2 in whatever state the preceding kernel left it, exactly as in the step.

Algorithmic work, no credit for lane padding:
  conv fwd / dgrad / wgrad : 2 * B*Ho*Wo * Cin*Cout*k*k FLOP;  bytes = s*(x + y) (+ packed weights), logical channels
  BN fwd (stats+apply)     : (3 [+1 residual]) * s B / element;  BN bwd: (5 [+1 residual gradient]) * s B / element
  fuse / copy / layout     : every source read once + destination written once
"""
import collections

import torch

from . import native as N

_ESZ = {0: 4, 1: 2}


def _conv_work(g, passname):
    cin, cout = getattr(g, "Cin", g.Cin_p), getattr(g, "Cout", g.Cout_p)
    flop = 2.0 * g.B * g.Ho * g.Wo * cin * cout * g.k * g.k
    return flop, g.B * g.H * g.W * cin, g.B * g.Ho * g.Wo * cout, g.k * g.k * cin * cout


def _geom_of(args):
    for a in args:
        o = getattr(a, "_obj", None)
        if isinstance(o, N.ConvGeom):
            return o
    return None


def describe(name, args):
    """(kernel label, shape label, algorithmic FLOP, algorithmic bytes) of one C-ABI call."""
    g = _geom_of(args)
    if g is not None and name.startswith("vae2_conv2d"):
        code = {"vae2_conv2d_fwd": lambda a: a[4], "vae2_conv2d_dgrad": lambda a: a[3], "vae2_conv2d_wgrad": lambda a: a[3],
                "vae2_conv2d_wgrad_tc": lambda a: 1, "vae2_conv2d_wgrad_f32x2": lambda a: 0}[name](args)
        s = _ESZ[code]
        flop, nx, ny, nw = _conv_work(g, name)
        by = s * (nx + ny) + (4 if "wgrad" in name else s) * nw
        what = {"vae2_conv2d_fwd": "fwd", "vae2_conv2d_dgrad": "dgrad", "vae2_conv2d_wgrad": "wgrad",
                "vae2_conv2d_wgrad_tc": "wgrad", "vae2_conv2d_wgrad_f32x2": "wgrad"}[name]
        shape = "%s %d->%d k%d s%d @%dx%d B=%d" % (what, getattr(g, "Cin", g.Cin_p), getattr(g, "Cout", g.Cout_p), g.k,
                                                   g.stride, g.H, g.W, g.B)
        kern = N.lib().vae2_last_kernel().decode() or name
        if name == "vae2_conv2d_wgrad_tc":
            kern = kern + "+wgrad_reduce_kernel"
        if name == "vae2_conv2d_wgrad_f32x2":
            dual = " dual" in kern
            kern = "split_planes+" + ("" if dual else "3x ") + kern.replace(" (f32x2 planes)", "") + "+reduce (fp32 wgrad)"
        return kern, shape, flop, float(by)
    if name in ("vae2_bn_fwd_fused", "vae2_bn_fwd_fused_groups"):
        code, npix, C_ = args[4], args[5], args[6]
        G = args[23] if name.endswith("groups") else 1
        res = args[1] is not None
        return ("bn_fwd_fused_kernel", "C=%d npix=%d%s%s" % (C_, npix, " +res" if res else "", " x%d groups" % G if G > 1 else ""),
                0.0, float(_ESZ[code] * npix * G * C_ * (3 + res)))
    if name in ("vae2_bn_bwd_fused", "vae2_bn_bwd_fused_groups"):
        code, npix, C_ = args[6], args[7], args[8]
        G = args[27] if name.endswith("groups") else 1
        dres = args[4] is not None
        relu = args[24]
        return ("bn_bwd_fused_kernel", "C=%d npix=%d relu=%d%s%s" % (C_, npix, relu, " +dres" if dres else "",
                                                                     " x%d groups" % G if G > 1 else ""), 0.0,
                float(_ESZ[code] * npix * G * C_ * (5 + dres + (1 if relu == 1 else 0))))
    if name in ("vae2_bn_stats", "vae2_bn_apply", "vae2_bn_bwd_reduce", "vae2_bn_bwd_elemt"):
        idx = {"vae2_bn_stats": (3, 4, 5, 1), "vae2_bn_apply": (3, 4, 5, 2), "vae2_bn_bwd_reduce": (5, 6, 7, 2),
               "vae2_bn_bwd_elemt": (5, 6, 7, 3)}[name]
        code, npix, Cp, passes = args[idx[0]], args[idx[1]], args[idx[2]], idx[3]
        return name[5:] + "_kernel", "Cp=%d npix=%d" % (Cp, npix), 0.0, float(_ESZ[code] * npix * Cp * passes)
    if name == "vae2_fuse_sum":
        srcs, n, code, B, H, W, Cp = args[0], args[1], args[3], args[4], args[5], args[6], args[7]
        el = B * H * W * Cp + sum(B * srcs[i].H * srcs[i].W * Cp for i in range(n))
        return "fuse_sum_kernel", "n=%d Cp=%d @%dx%d" % (n, Cp, H, W), 0.0, float(_ESZ[code] * el)
    if name == "vae2_fuse_bwd_up":
        code, B, H, W, Hs, Ws, Cp = args[3], args[4], args[5], args[6], args[7], args[8], args[9]
        return "fuse_bwd_up_kernel", "Cp=%d %dx%d->%dx%d" % (Cp, H, W, Hs, Ws), 0.0, float(_ESZ[code] * B * Cp * (2 * H * W + Hs * Ws))
    if name == "vae2_fuse_bwd_same":
        n, code, npix, Cp = args[3], args[4], args[5], args[6]
        return "fuse_bwd_same_kernel", "n=%d Cp=%d npix=%d" % (n, Cp, npix), 0.0, float(_ESZ[code] * npix * Cp * (2 + n))
    if name == "vae2_slice_copy":
        code, npix, Cp = args[2], args[3], args[4]
        return "slice_copy_kernel", "Cp=%d npix=%d" % (Cp, npix), 0.0, float(2 * _ESZ[code] * npix * Cp)
    if name in ("vae2_nchw_to_act", "vae2_act_to_nchw"):
        code, B = args[2], args[3]
        C_, H, W = (args[4], args[6], args[7]) if name == "vae2_nchw_to_act" else (args[4], args[5], args[6])
        return name[5:] + "_kernel", "C=%d @%dx%d" % (C_, H, W), 0.0, float((4 + _ESZ[code]) * B * C_ * H * W)
    if name == "vae2_bias_grad":
        code, npix, C_ = args[2], args[3], args[4]
        return "bias_grad_kernel", "C=%d npix=%d" % (C_, npix), 0.0, float(_ESZ[code] * npix * C_)
    return name[5:], "", 0.0, 0.0


class _TimedCaller:
    SPIN_EVERY = 128          # launches per device-side spin
    SPIN_CYCLES = 40_000_000  # ~20 ms at 1.9 GHz: well beyond what the host needs to enqueue SPIN_EVERY launches (~30 us each
                              # with the two event records; cooperative launches more) -- a starved GPU would time host latency

    def __init__(self, records):
        self.records, self.n = records, 0

    def __getattr__(self, name):
        fn = getattr(N.lib(), name)
        recs = self.records

        def wrapped(*args):
            N.COUNTERS["native_calls"] += 1
            if self.n % self.SPIN_EVERY == 0:
                torch.cuda._sleep(self.SPIN_CYCLES)
            self.n += 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(fn(*args), name)
            e1.record()
            recs.append((name, describe(name, args), e0, e1))
        setattr(self, name, wrapped)
        return wrapped


class KernelProfile:
    """with KernelProfile() as kp: <one eager training step>;  kp.rows() -> per (kernel, shape) totals."""

    def __enter__(self):
        self.records = []
        self._saved = N.call
        N.call = _TimedCaller(self.records)
        return self

    def __exit__(self, *exc):
        N.call = self._saved
        torch.cuda.synchronize()
        return False

    def rows(self):
        agg = collections.OrderedDict()
        for name, (kern, shape, flop, by), e0, e1 in self.records:
            a = agg.setdefault((kern, shape), {"kernel": kern, "shape": shape, "ms": 0.0, "n": 0, "flop": 0.0, "bytes": 0.0})
            a["ms"] += e0.elapsed_time(e1)
            a["n"] += 1
            a["flop"] += flop
            a["bytes"] += by
        return sorted(agg.values(), key=lambda r: -r["ms"])

    def by_kernel(self):
        fam = collections.OrderedDict()
        for r in self.rows():
            f = fam.setdefault(r["kernel"], {"kernel": r["kernel"], "ms": 0.0, "n": 0, "flop": 0.0, "bytes": 0.0, "top_shape": r})
            for k in ("ms", "n", "flop", "bytes"):
                f[k] += r[k]
        return sorted(fam.values(), key=lambda r: -r["ms"])
