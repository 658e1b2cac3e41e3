"""Per-call-site time table of one training iteration: every C-ABI launch is bracketed by CUDA events and
grouped by (entry point, conv geometry), so the convolution time can be read per layer shape."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
import bench
from config import load_config
from _engine_loader import engine

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
workload = sys.argv[2] if len(sys.argv) > 2 else "w18_256x512"
E = engine(); E.set_precision(prec); E.use_cuda_graphs(False)
yaml_name, H, W, B, _ = bench.WORKLOADS[workload]
if isinstance(B, dict):
    B = B[prec]
if len(sys.argv) > 3:
    B = int(sys.argv[3])
cfg = load_config(os.path.join(ROOT, "experiments", "vae2", yaml_name))
dev = torch.device("cuda:0")
g, d, og, od = bench.build_models(cfg, dev, 1, 0)
xt = torch.randn(B, 9, H, W, device=dev); x2t = xt + 0.1 * torch.randn_like(xt); x3t = x2t + 0.1 * torch.randn_like(xt)
for _ in range(2):
    bench.train_step(g, d, og, od, xt, x2t, x3t)
torch.cuda.synchronize()

N = E.native
records = []


def label(name, args):
    for a in args:
        o = getattr(a, "_obj", None)
        if isinstance(o, N.ConvGeom):
            return "%s  %dx%d %d->%d k%d s%d" % (name, o.H, o.W, o.Cin_p, o.Cout_p, o.k, o.stride)
    if name.startswith(("vae2_bn", "vae2_fuse")):
        ints = [a for a in args if isinstance(a, int) and not isinstance(a, bool) and 0 <= a < (1 << 24)]
        return name + "  " + ",".join(str(i) for i in ints[:6])
    return name


class TimedCaller:
    def __getattr__(self, name):
        fn = getattr(N.lib(), name)

        def wrapped(*args):
            N.COUNTERS["native_calls"] += 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(fn(*args), name)
            e1.record()
            records.append((label(name, args), e0, e1))
        setattr(self, name, wrapped)
        return wrapped


N.call = TimedCaller()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
bench.train_step(g, d, og, od, xt, x2t, x3t)
t1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for lab, e0, e1 in records:
    a = agg[lab]
    a[0] += e0.elapsed_time(e1); a[1] += 1
rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
tot = sum(r[0] for r in rows)
print("precision %s workload %s B=%d: eager step %.1f ms, %.1f ms inside %d timed C-ABI calls" %
      (prec, workload, B, t0.elapsed_time(t1), tot, len(records)))
for t, c, k in rows[:70]:
    print("%9.2f ms %5.1f%% n=%5d avg=%8.1f us  %s" % (t, 100 * t / tot, c, 1e3 * t / c, k))
