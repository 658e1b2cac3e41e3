"""Run the reference's UNMODIFIED training loop (lib/core/function.py::adversarial_train, :443-512) for a few
iterations  --  TEST INFRASTRUCTURE (launched as a subprocess by tests/test_dropin.py; not collected by pytest).

    python tests/dropin_driver.py --side mirror    --ref <reference root> --device cuda:0 --out a.json
    python tests/dropin_driver.py --side reference --ref <reference root> --device cpu    --out b.json

side=mirror   : sys.path = [stubs, vae-2_b200/lib, <ref>/lib]  -- the ONLY difference to the reference's own
                tools/_init_paths.py is that the drop-in tree comes first.  models / utils.utils / core.criterion /
                config resolve in the mirror, core.function (the loop under test) and everything else in the reference.
side=reference: sys.path = [stubs, <ref>/lib]                  -- the reference's own modules on the CPU.
Both sides are built exactly as tools/train.py:204-261 builds them (wrappers, identity asserts, Adam with the 'D_model'
name filter), get the same procedurally filled weights, the same clips and the same injected eps, and dump every
scalar the loop logged plus a few parameter norms after the last optimizer step.
"""
import argparse
import importlib.util
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cfg(name, opts=None):
    """The mirror's yacs-free config implementation, by file location (the reference's needs yacs)."""
    d = os.path.join(ROOT, "vae-2_b200", "lib", "config")
    spec = importlib.util.spec_from_file_location("vae2_cfg", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vae2_cfg"] = mod
    spec.loader.exec_module(mod)
    return mod.load_config(os.path.join(ROOT, "experiments", "vae2", name), opts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", required=True, choices=["mirror", "reference"])
    ap.add_argument("--ref", required=True)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--out", required=True)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--yaml", default="vae2_hrnet_tiny_32x64.yaml")
    ap.add_argument("--size", default="2,32,64")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--graphs", type=int, default=0)
    args = ap.parse_args()

    import numpy as np
    np.int = int                                     # numpy >= 1.24 removed it; enc_hrnet.py:321 (reference) uses it
    paths = [os.path.join(ROOT, "tests", "stubs")]
    if args.side == "mirror":
        paths.append(os.path.join(ROOT, "vae-2_b200", "lib"))
    paths.append(os.path.join(args.ref, "lib"))
    sys.path[:0] = paths
    sys.path.append(ROOT)
    import torch
    from oracle import vae2_oracle as O             # test infrastructure: deterministic weights / inputs / eps

    # --- tools/train.py:27-35, the reference driver's imports, verbatim in spirit -------------------------------------
    import models                                    # noqa: F401
    import models.enc_hrnet as enc_hrnet
    from core.criterion import L1Loss, KLLoss, lsgan_adversarial_loss
    from core.function import adversarial_train
    from utils.utils import FullModel_encdec, FullModel_D
    import core.function as F_
    import utils.utils as U_
    where = {"core.function": F_.__file__, "utils.utils": U_.__file__, "models.enc_hrnet": enc_hrnet.__file__}
    if args.side == "mirror":
        assert os.path.abspath(args.ref) in os.path.abspath(F_.__file__), F_.__file__
        assert "vae-2_b200" in U_.__file__ and "vae-2_b200" in enc_hrnet.__file__, where
        from _engine_loader import engine
        E = engine()
        E.set_precision(args.precision)
        E.use_cuda_graphs(bool(args.graphs))

    cfg = load_cfg(args.yaml, ["PRINT_FREQ", "1"])
    device = torch.device(args.device)
    # --- tools/train.py:204-261 ---------------------------------------------------------------------------------------
    encz_model, encdec_model = enc_hrnet.get_encz_model(cfg), enc_hrnet.get_encdec_model(cfg)
    D_model_sequence, D_model_frame = enc_hrnet.get_D_sequence_model(cfg), enc_hrnet.get_D_frame_model(cfg)
    criterion_gan = lsgan_adversarial_loss()
    model_encdec = FullModel_encdec(encz_model=encz_model, encdec_model=encdec_model, D_model_sequence=D_model_sequence,
                                    D_model_frame=D_model_frame, criterion_recon=L1Loss(), criterion_KL=KLLoss(),
                                    criterion_gan=criterion_gan, x1recon_lambda=cfg.TRAIN.X1RECON_LAMBDA,
                                    x2recon_lambda=cfg.TRAIN.X2RECON_LAMBDA, x3recon_lambda=cfg.TRAIN.X3RECON_LAMBDA,
                                    gan_lambda=cfg.TRAIN.GAN_LAMBDA)
    model_D = FullModel_D(D_model_sequence=D_model_sequence, D_model_frame=D_model_frame, criterion_gan=criterion_gan)
    assert model_encdec.D_model_sequence is model_D.D_model_sequence, "Unexpected behavior."
    O.fill_state_dict(model_encdec.state_dict(), seed_tag="dropin", mode="trained")
    model_encdec, model_D = model_encdec.to(device), model_D.to(device)
    optimizer_encdec = torch.optim.Adam([{"params": [p for n, p in model_encdec.named_parameters()
                                                     if p.requires_grad and "D_model" not in n]}], lr=cfg.TRAIN.LR)
    optimizer_D = torch.optim.Adam([{"params": [p for n, p in model_D.named_parameters()
                                                if p.requires_grad and "D_model" in n]}], lr=cfg.TRAIN.LR)

    B, H, W = (int(v) for v in args.size.split(","))
    Z = cfg.MODEL.EXTRA.Z_DIM
    loader, queue = [], []
    for it in range(args.iters):
        tag = "dropin:%d" % it
        loader.append((list(O.make_clips(tag, B, H, W)), ["synthetic_clip"]))
        eps_z, code = O.make_eps(tag, B, Z, H, W)
        queue += eps_z + [code]
    orig = torch.randn

    def fake_randn(*size, **kw):                      # utils.py:89-93 (4 maps), then enc_hrnet.py:456 (code)
        t = queue.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(t.shape) == shape, (tuple(t.shape), shape)
        return t.clone().to(kw.get("device") or "cpu")
    torch.randn = fake_randn

    from tensorboardX import SummaryWriter
    writer = SummaryWriter()
    writer_dict = {"writer": writer, "train_global_steps": 0, "valid_global_steps": 0}
    outdir = tempfile.mkdtemp(prefix="vae2_dropin_")
    try:
        adversarial_train(cfg, 0, 1, args.iters, cfg.TRAIN.LR, args.iters, loader, optimizer_encdec, optimizer_D,
                          model_encdec, model_D, writer_dict, device, outdir, use_multiplier=False,
                          is_baseline=cfg.MODEL.EXTRA.IS_BASELINE, baseline_mode=cfg.MODEL.EXTRA.BASELINE_MODE)
    finally:
        torch.randn = orig
    assert not queue, "the loop consumed fewer randn draws than injected"
    sd = model_encdec.state_dict()
    keys = ["encz_model.conv1.weight", "encdec_model.decf_last_layer_2.3.bias", "D_model_frame.last_layer.0.weight",
            "encdec_model.stage3.0.fuse_layers.2.0.1.0.weight", "encz_model.bn1.running_mean",
            "D_model_sequence.bn2.running_var"]
    # the reference's checkpoint format (tools/train.py:317-330), written by whichever side this is
    ckpt = os.path.join(outdir, "checkpoint_encdec.pth.tar")
    torch.save({"epoch": 1, "state_dict": model_encdec.state_dict(), "optimizer_encdec": optimizer_encdec.state_dict()}, ckpt)
    pngs = sorted(f for _, _, fs in os.walk(os.path.join(outdir, "vis")) for f in fs)
    json.dump({"where": where, "scalars": writer.scalars, "global_steps": writer_dict["train_global_steps"],
               "params": {k: float(sd[k].double().norm()) for k in keys},
               "nbt": int(sd["D_model_frame.bn1.num_batches_tracked"]), "pngs": len(pngs), "ckpt": ckpt},
              open(args.out, "w"))


if __name__ == "__main__":
    main()
