"""Diagnostic (not a test): per-tensor errors of the CUDA path vs the fp32 and fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from helpers import *

name, wmode = sys.argv[1], sys.argv[2]
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
E.set_precision(prec)
gold = golden(name); cfg = cfg_of(str(gold["cfg"]))
B, H, W, Z, xt, x2t, x3t, eps_z, code = case_inputs(name, gold)
g, d = build_product(cfg)
O.fill_state_dict(g.state_dict(), name, wmode)
sd = {k: v.clone() for k, v in g.state_dict().items()}

def oracle(dtype):
    s = {k: (v.detach().clone().to(dtype).requires_grad_("running" not in k) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    c = lambda t: t.to(dtype)
    losses, a, b, cc = O.full_encdec_forward(s, cfg, c(xt), c(x2t), c(x3t), [c(e) for e in eps_z], c(code))
    losses[0].backward()
    return s, [float(l) for l in losses], (a.detach(), b.detach(), cc.detach())

s32, l32, p32 = oracle(torch.float32)
s64, l64, p64 = oracle(torch.float64)
dev = "cuda:0"
g = g.to(dev).train()
with RandnQueue([code]):
    losses, x1p, x2p, x3p = g(xt=xt.to(dev), x2t=x2t.to(dev), x3t=x3t.to(dev), multiplier=1.0, eps=[e.to(dev) for e in eps_z])
losses[0].backward()
got = [float(l) for l in losses]
print("case", name, prec)
print("losses  mine-vs-64", np.abs(np.array(got) - np.array(l64)) / np.abs(np.array(l64)))
print("losses  o32-vs-64 ", np.abs(np.array(l32) - np.array(l64)) / np.abs(np.array(l64)))
for nm, m, a32, a64 in zip(("x1p", "x2p", "x3p"), (x1p, x2p, x3p), p32, p64):
    print("%s mine-vs-64 %.2e  mine-vs-o32 %.2e  o32-vs-64 %.2e  mine-vs-gold %.2e" % (nm, rel_err(m, a64), rel_err(m, a32), rel_err(a32, a64), rel_err(m, gold[nm])))
rows = []
for k, p in g.named_parameters():
    r64 = s64[k].grad
    if r64 is None or float(r64.norm()) < 1e-12:
        continue
    rows.append((rel_err(p.grad, r64), rel_err(s32[k].grad, r64), float(r64.norm()), k))
rows.sort(reverse=True)
print("worst param grads (mine-vs-64, o32-vs-64, |g64|, name):")
for r in rows[:25]:
    print("  %.2e  %.2e  %.3e  %s" % r)
import statistics
print("median mine-vs-64 %.2e median o32-vs-64 %.2e n=%d" % (statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows), len(rows)))
