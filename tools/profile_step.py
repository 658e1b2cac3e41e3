"""Per-kernel time table of one training iteration (torch.profiler / CUPTI), eager launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
import bench
from config import load_config
from _engine_loader import engine

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.train_step(g, d, og, od, xt, x2t, x3t)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = e.cuda_time_total
    if t > 0:
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("precision %s workload %s: total device time %.1f ms over %d kernel launches" % (prec, workload, tot / 1e3, sum(r[1] for r in rows)))
for t, c, k in rows[:28]:
    print("%9.2f ms %5.1f%% n=%6d avg=%8.1f us  %s" % (t / 1e3, 100 * t / tot, c, t / c, k[:95]))
