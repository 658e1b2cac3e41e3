"""Kernel durations of one GRAPH-REPLAYED training iteration, grouped by (kernel, grid size): CUPTI activity
records via torch.profiler's chrome trace, so the times are the in-situ ones (warm L2, no launch gaps)."""
import os, sys, json, collections, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch
import bench
from config import load_config
from _engine_loader import engine

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
workload = sys.argv[2] if len(sys.argv) > 2 else "w18_256x512"
graphs = (sys.argv[3] if len(sys.argv) > 3 else "graphs") == "graphs"
E = engine(); E.set_precision(prec); E.use_cuda_graphs(graphs)
yaml_name, H, W, B, _ = bench.WORKLOADS[workload]
if isinstance(B, dict):
    B = B[prec]
cfg = load_config(os.path.join(ROOT, "experiments", "vae2", yaml_name))
dev = torch.device("cuda:0")
g, d, og, od = bench.build_models(cfg, dev, 1, 0)
xt = torch.randn(B, 9, H, W, device=dev); x2t = xt + 0.1 * torch.randn_like(xt); x3t = x2t + 0.1 * torch.randn_like(xt)
for _ in range(3):
    bench.train_step(g, d, og, od, xt, x2t, x3t)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.train_step(g, d, og, od, xt, x2t, x3t)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "vae2_trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
agg = collections.defaultdict(lambda: [0.0, 0])
t_min, t_max = 1e30, 0
for e in ev:
    if e.get("cat") != "kernel":
        continue
    a = e.get("args", {})
    name = e["name"].split("(")[0].replace("void ", "").replace("vae2::", "")
    key = (name[:60], tuple(a.get("grid", [])), a.get("registers per thread"))
    agg[key][0] += e["dur"]; agg[key][1] += 1
    t_min, t_max = min(t_min, e["ts"]), max(t_max, e["ts"] + e["dur"])
rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
tot = sum(r[0] for r in rows)
print("precision %s workload %s B=%d graphs=%s: kernels %.1f ms busy, %.1f ms first-to-last, %d launches" %
      (prec, workload, B, graphs, tot / 1e3, (t_max - t_min) / 1e3, sum(r[1] for r in rows)))
for t, c, k in rows[:int(os.environ.get("TOP", "60"))]:
    print("%9.2f ms %5.1f%% n=%5d avg=%8.1f us  %-60s grid=%s regs=%s" % (t / 1e3, 100 * t / tot, c, t / c, k[0], k[1], k[2]))
