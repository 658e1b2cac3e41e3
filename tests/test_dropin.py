"""The drop-in boundary, exercised the way a reference user would use it:

* the import lines of the reference's tools/train.py (:27-35) resolve with the drop-in tree FIRST on sys.path and the
  reference's lib/ behind it (mirror modules for the hot path, the reference's own control plane for the rest);
* the reference's UNMODIFIED lib/core/function.py::adversarial_train drives the mirror for two iterations on the GPU and
  logs the same losses as it does when it drives the reference's own modules on the CPU
  (tests/golden/dropin_tiny_reference.json, written by tests/dropin_driver.py --side reference);
* checkpoints written by the reference modules load into the mirror (strict) and vice versa, and the ImageNet-pretrained
  remap (enc_hrnet.py:761-785) gives the same tensors.

A reference tree is needed: /root/reference in the build container, oracle/_ref (oracle/make_ref.py) on the GPU box.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import O, ROOT, GOLD, build_product, cfg_of, log_err

STUBS = os.path.join(ROOT, "tests", "stubs")
MIRROR = os.path.join(ROOT, "vae-2_b200", "lib")


def _ref_root():
    for p in (os.environ.get("VAE2_REFERENCE"), os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if p and os.path.isfile(os.path.join(p, "lib", "core", "function.py")):
            return p
    return None


REF = _ref_root()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference tree (run oracle/make_ref.py in the build container)")


def _run(code_or_args, script=False, timeout=900):
    cmd = [sys.executable] + (code_or_args if script else ["-c", code_or_args])
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
    return r.stdout


@needs_ref
def test_reference_train_py_imports_resolve_over_the_mirror():
    code = r'''
import sys, json, numpy as np
np.int = int
sys.path[:0] = [%r, %r, %r]
# tools/train.py:25-35
from tensorboardX import SummaryWriter
import models
import datasets
from config import config
from config import update_config
from core.criterion import *
from core.function import train, validate, adversarial_train
from utils.modelsummary import get_model_summary
from utils.utils import create_logger, FullModel, get_rank, FullModel_all, FullModel_encdec, FullModel_D
# tools/inference.py / tools/test.py extras
from core.function import inference, testval, test
import models.seg_hrnet, utils.metric
import core.function, core.criterion, utils.utils, models.enc_hrnet
print(json.dumps({"function": core.function.__file__, "criterion": core.criterion.__file__,
                  "utils": utils.utils.__file__, "enc_hrnet": models.enc_hrnet.__file__,
                  "seg_hrnet": models.seg_hrnet.__file__, "datasets": datasets.__file__,
                  "FullModel": FullModel.__module__, "FullModel_encdec": FullModel_encdec.__module__,
                  "CrossEntropy": CrossEntropy.__module__, "L1Loss": L1Loss.__module__,
                  "has": [n for n in ("OhemCrossEntropy", "KLLoss", "lsgan_adversarial_loss", "PSNR") if n in globals()]}))
''' % (STUBS, MIRROR, os.path.join(REF, "lib"))
    out = json.loads(_run(code).strip().splitlines()[-1])
    for k in ("criterion", "utils", "enc_hrnet"):
        assert out[k].startswith(MIRROR), (k, out[k])                 # the hot path is the mirror's
    for k in ("function", "seg_hrnet", "datasets"):
        assert out[k].startswith(REF), (k, out[k])                    # the control plane is the reference's own
    assert out["FullModel_encdec"] == "utils.utils" and out["L1Loss"] == "core.criterion"
    assert out["FullModel"].startswith("_vae2_ref.") and out["CrossEntropy"].startswith("_vae2_ref.")
    assert len(out["has"]) == 4


@needs_ref
def test_reference_checkpoints_and_pretrained_remap_interoperate(tmp_path):
    """f4: on-disk formats either side of the path (tools/train.py:270-290, 317-348; enc_hrnet.py:761-785)."""
    ck = str(tmp_path / "checkpoint_encdec.pth.tar")
    pre = str(tmp_path / "hrnet_imagenet_like.pth")
    out = str(tmp_path / "ref_after_pretrained.pth")
    code = r'''
import sys, os, numpy as np, torch, importlib.util
np.int = int
ROOT = %r
sys.path[:0] = [%r]
sys.path.append(ROOT)
from oracle import vae2_oracle as O
d = os.path.join(ROOT, "vae-2_b200", "lib", "config")
spec = importlib.util.spec_from_file_location("vae2_cfg", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
mod = importlib.util.module_from_spec(spec); sys.modules["vae2_cfg"] = mod; spec.loader.exec_module(mod)
cfg = mod.load_config(os.path.join(ROOT, "experiments", "vae2", "vae2_hrnet_tiny_32x64.yaml"))
import models.enc_hrnet as M, utils.utils as U, core.criterion as Cr
nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss())
O.fill_state_dict(g.state_dict(), "ckpt", "trained")
opt = torch.optim.Adam([p for n, p in g.named_parameters() if "D_model" not in n], lr=1e-4)
torch.save({"epoch": 3, "state_dict": g.state_dict(), "optimizer_encdec": opt.state_dict()}, %r)      # train.py:320-324
# an ImageNet-HRNet-like blob: 'model.'-prefixed trunk weights with a 3-channel conv1 (enc_hrnet.py:761-785)
src = M.get_D_frame_model(cfg)
O.fill_state_dict(src.state_dict(), "pre", "trained")
torch.save({"model." + k: v for k, v in src.state_dict().items()}, %r)
res = {}
for seed in (1, 2):      # init_weights draws N(0, 1e-3) first: tensors equal under both seeds are the remapped ones
    for name, fn in (("encdec", M.get_encdec_model), ("encz", M.get_encz_model), ("dseq", M.get_D_sequence_model)):
        torch.manual_seed(seed)
        m = fn(cfg)
        m.init_weights(%r)
        res["%%s:%%d" %% (name, seed)] = m.state_dict()
torch.save(res, %r)
''' % (ROOT, os.path.join(REF, "lib"), ck, pre, pre, out)
    _run(code)
    cfg = cfg_of("vae2_hrnet_tiny_32x64.yaml")
    g, d = build_product(cfg)
    blob = torch.load(ck, map_location="cpu")
    g.load_state_dict(blob["state_dict"], strict=True)                  # train.py:282 (model.module.load_state_dict)
    want = {k: v.clone() for k, v in g.state_dict().items()}
    O.fill_state_dict(want, "ckpt", "trained")
    for k, v in g.state_dict().items():
        assert torch.equal(v, want[k]), k
    # optimizer state of the reference checkpoint fits an optimizer built over the mirror's parameters (same order)
    opt = torch.optim.Adam([p for n, p in g.named_parameters() if "D_model" not in n], lr=1e-4)
    opt.load_state_dict(blob["optimizer_encdec"])
    assert blob["epoch"] == 3
    # pretrained remap: same tensors as the reference's init_weights(pretrained) produces
    import models.enc_hrnet as M
    ref_after = torch.load(out, map_location="cpu")
    for name, fn in (("encdec", M.get_encdec_model), ("encz", M.get_encz_model), ("dseq", M.get_D_sequence_model)):
        m = fn(cfg)
        m.init_weights(pre)
        mine = m.state_dict()
        a, b = ref_after[name + ":1"], ref_after[name + ":2"]
        assert sorted(mine) == sorted(a)
        fixed = [k for k in a if torch.equal(a[k], b[k]) and a[k].is_floating_point() and a[k].dim() == 4]
        assert len(fixed) > 50, (name, len(fixed))         # conv weights that came from the pretrained blob
        for k in fixed:
            assert torch.equal(mine[k], a[k]), (name, k)
        loaded_rand = [k for k in a if a[k].dim() == 4 and k not in fixed]
        assert all(mine[k].shape == a[k].shape for k in loaded_rand) and loaded_rand, name   # e.g. transition3_e, heads


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [0, 1], ids=["eager", "graphs"])
def test_reference_adversarial_train_drives_the_mirror(tmp_path, graphs):
    out = str(tmp_path / "mirror.json")
    _run([os.path.join(ROOT, "tests", "dropin_driver.py"), "--side", "mirror", "--ref", REF, "--device", "cuda:0",
          "--out", out, "--graphs", str(graphs)], script=True)
    got = json.load(open(out))
    ref = json.load(open(os.path.join(GOLD, "dropin_tiny_reference.json")))
    assert got["where"]["core.function"].startswith(REF) and "vae-2_b200" in got["where"]["utils.utils"]
    assert got["nbt"] == ref["nbt"] == 18 and got["pngs"] == ref["pngs"] == 18
    gs, rs = got["scalars"], ref["scalars"]
    assert [s[0] for s in gs] == [s[0] for s in rs] and len(gs) == 20
    e = np.array([abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(gs, rs)])
    pe = {k: abs(got["params"][k] - v) / v for k, v in ref["params"].items()}
    log_err("dropin_adversarial_train_graphs%d" % graphs, iter0=e[:10].max(), iter1=e[10:].max(), params=pe)
    assert e[:10].max() < 1e-4, e[:10]        # first iteration: forward parity
    assert e[10:].max() < 5e-3, e[10:]        # second iteration: after one Adam step of G and D on both sides
    assert max(pe.values()) < 1e-4, pe        # parameter / running-stat norms after two optimizer steps
    # the checkpoint the loop's caller writes (tools/train.py:317-330) from the MIRROR loads into the reference modules
    code = r'''
import sys, os, numpy as np, torch, importlib.util
np.int = int
ROOT = %r
sys.path[:0] = [%r]
d = os.path.join(ROOT, "vae-2_b200", "lib", "config")
spec = importlib.util.spec_from_file_location("vae2_cfg", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
mod = importlib.util.module_from_spec(spec); sys.modules["vae2_cfg"] = mod; spec.loader.exec_module(mod)
cfg = mod.load_config(os.path.join(ROOT, "experiments", "vae2", "vae2_hrnet_tiny_32x64.yaml"))
import models.enc_hrnet as M, utils.utils as U, core.criterion as Cr
nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss())
blob = torch.load(%r, map_location="cpu")
g.load_state_dict(blob["state_dict"], strict=True)
opt = torch.optim.Adam([p for n, p in g.named_parameters() if "D_model" not in n], lr=1e-4)
opt.load_state_dict(blob["optimizer_encdec"])
print("loaded", len(blob["state_dict"]))
''' % (ROOT, os.path.join(REF, "lib"), got["ckpt"])
    assert "loaded" in _run(code)
