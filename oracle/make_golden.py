"""Generate tests/golden/*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The reference modules (lib/models/enc_hrnet.py, lib/models/toy_fc.py,
lib/utils/utils.py, lib/core/criterion.py) are imported as they are; the only
shims are ``np.int = int`` (removed from numpy; used at enc_hrnet.py:321,596,700)
and an attribute-dict config.  ``torch.randn`` is patched to pop pre-drawn eps in
the reference's call order (4 posterior maps, utils.py:92-93, then the encoder's
random code, enc_hrnet.py:456).  Weights/inputs/eps come from
oracle.vae2_oracle.det_* so tests can regenerate them bit-exactly without the
reference being present.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("VAE2_REFERENCE", "/root/reference")

from oracle import vae2_oracle as O  # noqa: E402


def load_cfg(name, opts=None):
    spec = importlib.util.spec_from_file_location(
        "vae2_cfg", os.path.join(ROOT, "vae-2_b200", "lib", "config", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "vae-2_b200", "lib", "config")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vae2_cfg"] = mod
    spec.loader.exec_module(mod)
    return mod.load_config(os.path.join(ROOT, "experiments", "vae2", name), opts)


def import_reference():
    np.int = int  # shim 1 (numpy>=1.24 removed np.int)
    sys.path.insert(0, os.path.join(REF, "lib"))
    import models.enc_hrnet as enc_hrnet      # noqa
    import models.toy_fc as toy_fc            # noqa
    import utils.utils as rutils              # noqa
    import core.criterion as rcrit            # noqa
    return enc_hrnet, toy_fc, rutils, rcrit


class RandnQueue:
    """Patch torch.randn so the reference consumes injected eps in call order."""

    def __init__(self, tensors):
        self.q = list(tensors)
        self.orig = torch.randn

    def __enter__(self):
        def fake(*size, **kw):
            t = self.q.pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.clone()
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig
        assert not self.q, "reference consumed fewer randn calls than injected"


def build_reference(cfg, enc_hrnet, rutils, rcrit):
    nets = [enc_hrnet.get_encz_model(cfg), enc_hrnet.get_encdec_model(cfg),
            enc_hrnet.get_D_sequence_model(cfg), enc_hrnet.get_D_frame_model(cfg)]
    g = rutils.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], rcrit.L1Loss(), rcrit.KLLoss(),
                                rcrit.lsgan_adversarial_loss(),
                                x1recon_lambda=cfg.TRAIN.X1RECON_LAMBDA, x2recon_lambda=cfg.TRAIN.X2RECON_LAMBDA,
                                x3recon_lambda=cfg.TRAIN.X3RECON_LAMBDA, gan_lambda=cfg.TRAIN.GAN_LAMBDA)
    d = rutils.FullModel_D(nets[2], nets[3], rcrit.lsgan_adversarial_loss())
    return g, d


def grad_summary(model):
    names, norms, sums = [], [], []
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(n)
        norms.append(float(p.grad.double().norm()))
        sums.append(float(p.grad.double().sum()))
    return np.array(names), np.array(norms), np.array(sums)


def hrnet_case(fname, cfg_name, B, H, W, wmode, enc_hrnet, rutils, rcrit, keep_grads=()):
    cfg = load_cfg(cfg_name)
    g, d = build_reference(cfg, enc_hrnet, rutils, rcrit)
    sd = g.state_dict()
    O.fill_state_dict(sd, seed_tag=fname, mode=wmode)
    g.load_state_dict(sd)
    sd0 = {k: v.clone() for k, v in sd.items()}   # state_dict() aliases the live buffers: keep a snapshot
    Z = cfg.MODEL.EXTRA.Z_DIM
    xt, x2t, x3t = O.make_clips(fname, B, H, W)
    eps_z, code = O.make_eps(fname, B, Z, H, W, hd_z=bool(cfg.MODEL.EXTRA.HD_Z))
    out = {"meta": np.array([B, H, W, Z]), "cfg": np.array(cfg_name), "wmode": np.array(wmode),
           "hd_z": np.array(int(bool(cfg.MODEL.EXTRA.HD_Z)))}

    # --- G step, training mode ---
    g.train()
    with RandnQueue(eps_z + [code]):
        losses, x1p, x2p, x3p = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0)
    g.zero_grad()
    losses[0].backward()
    out["g_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
    out["x1p"], out["x2p"], out["x3p"] = x1p.detach().numpy(), x2p.detach().numpy(), x3p.detach().numpy()
    n, nr, sm = grad_summary(g)
    out["g_grad_names"], out["g_grad_norms"], out["g_grad_sums"] = n, nr, sm
    for k in keep_grads:
        out["grad:" + k] = dict(g.named_parameters())[k].grad.numpy().copy()
    sd_after = g.state_dict()
    for k in ("encz_model.bn1.running_mean", "encz_model.bn1.running_var",
              "encdec_model.decf_bn2.running_mean", "D_model_frame.bn1.running_var",
              "D_model_frame.bn1.num_batches_tracked"):
        out["after:" + k] = sd_after[k].numpy().copy()

    # --- the same G step with the reference in float64: the exact-arithmetic yardstick that tells
    #     rounding noise of a deep fp32 net apart from real disagreement (stored rounded to fp32) ---
    g64, _ = build_reference(cfg, enc_hrnet, rutils, rcrit)
    g64.load_state_dict(sd0)
    g64 = g64.double().train()
    with RandnQueue([e.double() for e in eps_z] + [code.double()]):
        l64, a64, b64, c64 = g64(xt=xt.double(), x2t=x2t.double(), x3t=x3t.double(), multiplier=1.0)
    out["g_losses64"] = np.array([float(l) for l in l64], dtype=np.float64)
    out["x1p64"], out["x2p64"], out["x3p64"] = (t.detach().float().numpy() for t in (a64, b64, c64))
    del g64

    # --- D step, training mode (continues from the updated BN running stats) ---
    dl = d(x2t=x2t, x2t_predict=x2p.detach())
    d.zero_grad()
    dl[0].backward()
    out["d_losses"] = np.array([float(l) for l in dl], dtype=np.float64)
    n, nr, sm = grad_summary(d)
    out["d_grad_names"], out["d_grad_norms"], out["d_grad_sums"] = n, nr, sm

    # --- eval-mode prior sampling (inference path, function.py:125-136) ---
    g.load_state_dict(sd0)  # back to the pre-step running stats
    g.eval()
    with torch.no_grad(), RandnQueue(eps_z + [code]):
        losses, x1e, x2e, x3e = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0, sampling_mode="prior_sampling")
    out["eval_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
    out["eval_x1p"], out["eval_x2p"], out["eval_x3p"] = x1e.numpy(), x2e.numpy(), x3e.numpy()
    path = os.path.join(ROOT, "tests", "golden", fname + ".npz")
    np.savez_compressed(path, **out)
    print(fname, "g_losses", out["g_losses"], "d_losses", out["d_losses"], os.path.getsize(path) // 1024, "KiB")


SUB = 8   # spatial subsampling stride of the full-size fixtures (whole tensors would be 4.7 MB each)


def _summ(t):
    """What a full-size fixture keeps of a prediction: every 8th pixel, plus per-channel mean and L2 norm over ALL pixels."""
    t = t.detach()
    d = t.double()
    return (t[..., ::SUB, ::SUB].float().numpy().copy(), d.mean(dim=(0, 2, 3)).numpy().copy(),
            d.pow(2).sum(dim=(0, 2, 3)).sqrt().numpy().copy())


def hrnet_case_fullsize(fname, cfg_name, B, H, W, wmode, enc_hrnet, rutils, rcrit, backward=True):
    """BASELINE-size case (configs[1]: W18 256x512; configs[4]: W48 473x473): the G step of the unmodified reference in
    training mode (batch-statistics BN), fp32 and fp64.  Predictions are stored subsampled (see _summ); with
    ``backward`` the fp32 parameter-gradient norms are stored as well, otherwise the pass runs under no_grad."""
    import contextlib
    cfg = load_cfg(cfg_name)
    g, _ = build_reference(cfg, enc_hrnet, rutils, rcrit)
    sd = g.state_dict()
    O.fill_state_dict(sd, seed_tag=fname, mode=wmode)
    g.load_state_dict(sd)
    sd0 = {k: v.clone() for k, v in sd.items()}
    Z = cfg.MODEL.EXTRA.Z_DIM
    xt, x2t, x3t = O.make_clips(fname, B, H, W)
    eps_z, code = O.make_eps(fname, B, Z, H, W)
    out = {"meta": np.array([B, H, W, Z]), "cfg": np.array(cfg_name), "wmode": np.array(wmode), "sub": np.array(SUB)}
    g.train()
    with (contextlib.nullcontext() if backward else torch.no_grad()), RandnQueue(eps_z + [code]):
        losses, x1p, x2p, x3p = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0)
    out["g_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
    for k, t in (("x1p", x1p), ("x2p", x2p), ("x3p", x3p)):
        out[k], out[k + "_mean"], out[k + "_l2"] = _summ(t)
    if backward:
        g.zero_grad()
        losses[0].backward()
        n, nr, sm = grad_summary(g)
        out["g_grad_names"], out["g_grad_norms"], out["g_grad_sums"] = n, nr, sm
    sd_after = g.state_dict()
    for k in ("encz_model.bn1.running_mean", "encdec_model.decf_bn2.running_mean", "D_model_frame.bn1.running_var"):
        out["after:" + k] = sd_after[k].numpy().copy()
    del g, losses, x1p, x2p, x3p
    g64, _ = build_reference(cfg, enc_hrnet, rutils, rcrit)
    g64.load_state_dict(sd0)
    g64 = g64.double().train()
    with torch.no_grad(), RandnQueue([e.double() for e in eps_z] + [code.double()]):
        l64, a64, b64, c64 = g64(xt=xt.double(), x2t=x2t.double(), x3t=x3t.double(), multiplier=1.0)
    out["g_losses64"] = np.array([float(l) for l in l64], dtype=np.float64)
    for k, t in (("x1p64", a64), ("x2p64", b64), ("x3p64", c64)):
        out[k], out[k + "_mean"], out[k + "_l2"] = _summ(t)
    path = os.path.join(ROOT, "tests", "golden", fname + ".npz")
    np.savez_compressed(path, **out)
    print(fname, "g_losses", out["g_losses"], "g_losses64", out["g_losses64"], os.path.getsize(path) // 1024, "KiB")


def toy_case(fname, toy_fc, rutils, rcrit):
    cfg = load_cfg("vae2_hrnet_tiny_32x64.yaml")  # only MODEL.EXTRA.IS_BASELINE/BASELINE_MODE are read
    nets = [toy_fc.get_encz_model(cfg), toy_fc.get_encdec_model(cfg), toy_fc.get_D_model(cfg)]
    g = rutils.FullToyModel_encdec(nets[0], nets[1], nets[2], rcrit.L1Loss(), rcrit.KLLoss(),
                                   rcrit.lsgan_adversarial_loss(), 1.0, 0.1, 1.0, 1.0)
    sd = g.state_dict()
    O.fill_state_dict(sd, seed_tag=fname, mode="trained")
    g.load_state_dict(sd)
    B = 500  # tools/toy_example.py:104-113 batches of 500 alphas
    # sigmoid curves as adversarial_train._gen_toyexample_data builds them (function.py:448-462)
    alphas = np.arange(0.001, 10.001, 0.001)[:B * 20:20]
    xt_t = np.arange(-1.5, -0.5, 0.1)[:10]
    rs = np.random.RandomState(7)
    x2_t = np.stack([rs.uniform(-0.5 + i / 10.0, -0.5 + (i + 1) / 10.0, size=B) for i in range(10)], 1)
    x3_t = np.stack([rs.uniform(0.5 + i / 10.0, 0.5 + (i + 1) / 10.0, size=B) for i in range(10)], 1)
    sig = lambda a, t: 1.0 / (1.0 + np.exp(-a * t))
    xt = torch.tensor(sig(alphas[:, None], xt_t[None, :]), dtype=torch.float32)
    x2t = torch.tensor(sig(alphas[:, None], x2_t), dtype=torch.float32)
    x3t = torch.tensor(sig(alphas[:, None], x3_t), dtype=torch.float32)
    eps = O.det_normal(fname + ":eps", (B, 8))
    code = O.det_normal(fname + ":code", (B, 8))
    g.train()
    # call order in the reference: reparam randn (utils.py:206) then _encoder_forward's random code (toy_fc.py:110)
    with RandnQueue([eps, code]):
        losses, x1p, x2p, x3p = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=0.5)
    g.zero_grad()
    losses[0].backward()
    n, nr, sm = grad_summary(g)
    path = os.path.join(ROOT, "tests", "golden", fname + ".npz")
    np.savez_compressed(path, xt=xt.numpy(), x2t=x2t.numpy(), x3t=x3t.numpy(),
                        losses=np.array([float(l) for l in losses]), x1p=x1p.detach().numpy(),
                        x2p=x2p.detach().numpy(), x3p=x3p.detach().numpy(),
                        grad_names=n, grad_norms=nr, grad_sums=sm)
    print(fname, "losses", [float(l) for l in losses])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    enc_hrnet, toy_fc, rutils, rcrit = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "nohdz":     # non-HD_Z posterior head
        hrnet_case("tiny_b2_32x64_nohdz", "vae2_hrnet_tiny_32x64_nohdz.yaml", 2, 32, 64, "trained", enc_hrnet, rutils, rcrit)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "w18_full":   # BASELINE configs[1] size (about 2 min, 30 GB of host memory)
        hrnet_case_fullsize("w18_b1_256x512", "vae2_hrnet_w18_small_v2_256x512.yaml", 1, 256, 512, "trained",
                            enc_hrnet, rutils, rcrit, backward=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "w48_full":   # BASELINE configs[4] LIP size, forward only (about 15 min)
        hrnet_case_fullsize("w48_b1_473x473", "vae2_hrnet_w48_473x473.yaml", 1, 473, 473, "trained",
                            enc_hrnet, rutils, rcrit, backward=False)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "w48":       # regenerate only the W48 case
        hrnet_case("w48_b1_33x33", "vae2_hrnet_w48_473x473.yaml", 1, 33, 33, "trained", enc_hrnet, rutils, rcrit)
        return
    hrnet_case("tiny_b2_32x64", "vae2_hrnet_tiny_32x64.yaml", 2, 32, 64, "trained", enc_hrnet, rutils, rcrit,
               keep_grads=("encz_model.conv1.weight", "encdec_model.transition3_e.0.0.weight",
                           "encdec_model.decp_last_layer_2.3.bias", "encdec_model.stage3.0.fuse_layers.2.0.1.0.weight"))
    hrnet_case("tiny_b1_33x47", "vae2_hrnet_tiny_32x64.yaml", 1, 33, 47, "trained", enc_hrnet, rutils, rcrit)
    hrnet_case("tiny_b2_32x64_init", "vae2_hrnet_tiny_32x64.yaml", 2, 32, 64, "init", enc_hrnet, rutils, rcrit)
    hrnet_case("w18_b1_32x64", "vae2_hrnet_w18_small_v2_256x512.yaml", 1, 32, 64, "trained", enc_hrnet, rutils, rcrit)
    # HRNet-W48 (LIP / PASCAL-Context experiments, BASELINE configs[4]) at a small odd square size
    hrnet_case("w48_b1_33x33", "vae2_hrnet_w48_473x473.yaml", 1, 33, 33, "trained", enc_hrnet, rutils, rcrit)
    toy_case("toy_b500", toy_fc, rutils, rcrit)


if __name__ == "__main__":
    main()
