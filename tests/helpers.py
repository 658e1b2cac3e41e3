"""Shared test helpers: build the product modules, fill them like the golden generator did."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import vae2_oracle as O  # noqa: E402  (test infrastructure)

import models.enc_hrnet as M          # noqa: E402  (product, via vae-2_b200/lib on sys.path)
import models.toy_fc as T             # noqa: E402
import utils.utils as U               # noqa: E402
import core.criterion as Cr           # noqa: E402
from config import load_config        # noqa: E402
from _engine_loader import engine     # noqa: E402

E = engine()


def cfg_of(name):
    return load_config(os.path.join(ROOT, "experiments", "vae2", name))


def build_product(cfg):
    nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
    g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(),
                           Cr.lsgan_adversarial_loss(), cfg.TRAIN.X1RECON_LAMBDA, cfg.TRAIN.X2RECON_LAMBDA,
                           cfg.TRAIN.X3RECON_LAMBDA, cfg.TRAIN.GAN_LAMBDA)
    d = U.FullModel_D(nets[2], nets[3], Cr.lsgan_adversarial_loss())
    return g, d


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def case_inputs(name, gold):
    B, H, W, Z = [int(v) for v in gold["meta"]]
    xt, x2t, x3t = O.make_clips(name, B, H, W)
    hd = bool(int(gold["hd_z"])) if "hd_z" in gold.files else True
    eps_z, code = O.make_eps(name, B, Z, H, W, hd_z=hd)
    return B, H, W, Z, xt, x2t, x3t, eps_z, code


class RandnQueue:
    """Inject pre-drawn tensors into torch.randn calls (same trick as oracle/make_golden.py)."""

    def __init__(self, tensors):
        self.q = list(tensors)
        self.orig = torch.randn

    def __enter__(self):
        def fake(*size, **kw):
            t = self.q.pop(0)
            dev = kw.get("device")
            return t.clone().to(dev) if dev is not None else t.clone()
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.floating, np.integer)):
        return float(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    return v


def log_err(case, **vals):
    """Measured parity errors, one JSON line per case: printed (pytest -s / -rP shows it) and appended to
    gpurun_out/parity_errors.jsonl so a GPU run brings them back ('log measured errors per case')."""
    import json
    rec = {"case": case}
    rec.update(_plain(vals))
    line = json.dumps(rec)
    print("PARITY", line)
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.jsonl"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
