"""2-GPU check (run under torchrun, not collected by pytest): DDP + SyncBN on 2 ranks with one sample each must
reproduce the single-process B=2 golden G step of the reference (losses, predictions, averaged gradients)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import numpy as np
import torch
import torch.distributed as dist

from helpers import E, O, RandnQueue, build_product, case_inputs, cfg_of, golden, rel_err

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl")
P2P = "p2p" in sys.argv[2:]                           # SyncBN statistics over NVLink peer memory instead of NCCL (engine/peer.py)
if P2P:
    assert E.peer.enable(), "peer-memory SyncBN could not be enabled"
else:
    os.environ["VAE2_SYNCBN_P2P"] = "0"
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
E.set_precision(prec)
E.use_cuda_graphs(len(sys.argv) > 2 and sys.argv[2] == "graphs")
name = "tiny_b2_32x64"
gold = golden(name)
cfg = cfg_of(str(gold["cfg"]))
g, d = build_product(cfg)
O.fill_state_dict(g.state_dict(), seed_tag=name, mode="trained")
B, H, W, Z, xt, x2t, x3t, eps_z, code = case_inputs(name, gold)
assert B == world
g.skip_dead_D_grads = "skip" in sys.argv[2:]         # eager GAN terms with the discriminators frozen in the G step
g = torch.nn.SyncBatchNorm.convert_sync_batchnorm(g).to(dev).train()
d = torch.nn.SyncBatchNorm.convert_sync_batchnorm(d).to(dev).train()
dd = torch.nn.parallel.DistributedDataParallel(d, device_ids=[rank], find_unused_parameters=True)
gd = torch.nn.parallel.DistributedDataParallel(g, device_ids=[rank], find_unused_parameters=True)
if P2P:
    E.peer.serialize_ddp(dd)
    E.peer.serialize_ddp(gd)
sl = slice(rank, rank + 1)
with RandnQueue([code[sl]]):
    losses, x1p, x2p, x3p = gd(xt=xt[sl].to(dev), x2t=x2t[sl].to(dev), x3t=x3t[sl].to(dev), multiplier=1.0,
                               eps=[e[sl].to(dev) for e in eps_z])
losses[0].backward()
lv = torch.stack([l.detach().reshape(()) for l in losses])
dist.all_reduce(lv)
lv = (lv / world).cpu().numpy()
tol = 1e-4 if prec == "fp32" else 2e-2
ok = True
msgs = []
rel = np.abs(lv - gold["g_losses"]) / np.abs(gold["g_losses"])
msgs.append("mean-of-rank losses vs reference B=2: max rel err %.2e" % rel.max())
if prec == "fp32":
    ok &= bool(rel.max() < tol)
else:   # bf16 path: the ELBO (3 x L1 + KL) within 1e-2 (north star); GAN terms are reported only
    elbo, ref = lv[1:5].sum(), gold["g_losses"][1:5].sum()
    msgs.append("ELBO rel err %.2e" % (abs(elbo - ref) / abs(ref)))
    ok &= bool(abs(elbo - ref) / abs(ref) < 1e-2)
e2 = rel_err(x2p, gold["x2p"][sl])
msgs.append("x2p[rank] rel err %.2e" % e2)
ok &= e2 < (1e-4 if prec == "fp32" else 8e-2)
if prec == "fp32":
    norms = dict(zip(gold["g_grad_names"].tolist(), gold["g_grad_norms"]))
    en = np.array([abs(float(p.grad.double().norm()) - norms[k]) / norms[k] for k, p in g.named_parameters()
                   if norms[k] > 1e-9 and ".0.bias" not in k and p.grad is not None])
    msgs.append("DDP-averaged grad norms vs reference: median rel err %.2e, max %.2e" % (np.median(en), en.max()))
    ok &= bool(np.median(en) < 2e-2)
sd = g.state_dict()
e3 = rel_err(sd["encz_model.bn1.running_mean"], gold["after:encz_model.bn1.running_mean"])
msgs.append("synced running_mean rel err %.2e" % e3)
ok &= e3 < (1e-4 if prec == "fp32" else 5e-2)
# D step (stacked passes, eager backward) under DDP + SyncBN against the reference's single-process B=2 D losses / grad norms
x2_full = torch.from_numpy(gold["x2p"]).to(dev)
dl = dd(x2t=x2t[sl].to(dev), x2t_predict=x2_full[sl])
d.zero_grad()
dl[0].backward()
dv = torch.stack([l.detach().reshape(()) for l in dl])
dist.all_reduce(dv)
dv = (dv / world).cpu().numpy()
drel = np.abs(dv - gold["d_losses"]) / np.abs(gold["d_losses"])
msgs.append("D step: mean-of-rank losses vs reference B=2: max rel err %.2e" % drel.max())
ok &= bool(drel.max() < (2e-4 if prec == "fp32" else 3e-2))
if prec == "fp32":
    dn = dict(zip(gold["d_grad_names"].tolist(), gold["d_grad_norms"]))
    en = np.array([abs(float(p.grad.double().norm()) - dn[k]) / dn[k] for k, p in d.named_parameters()
                   if dn[k] > 1e-9 and not k.endswith("last_layer.0.bias")])
    msgs.append("D step: DDP-averaged grad norms vs reference: median rel err %.2e, max %.2e" % (np.median(en), en.max()))
    ok &= bool(np.median(en) < 3e-2)
plans = [p for m in g.modules() if hasattr(m, "_plans") for pool in m._plans().values() for p in pool]
msgs.append("SyncBN collectives fwd %d, peer-memory launches fwd %d bwd %d, for %d BNs" % (
    sum(p.n_collectives_fwd for p in plans), sum(getattr(p, "n_peer_bn_fwd", 0) for p in plans),
    sum(getattr(p, "n_peer_bn_bwd", 0) for p in plans), sum(1 for m in g.modules() if isinstance(m, torch.nn.SyncBatchNorm))))
if P2P:
    E.peer.check()
    # (a BN whose output slice is narrower than its input keeps the NCCL path: a handful in bf16)
    ok &= sum(p.n_collectives_fwd for p in plans) <= 4 and sum(getattr(p, "n_peer_bn_fwd", 0) for p in plans) > 300
if rank == 0:
    print("\n".join(msgs))
    print("DIST_SYNCBN_CHECK", "PASS" if ok else "FAIL", prec, flush=True)
dist.barrier()
os._exit(0 if ok else 1)
