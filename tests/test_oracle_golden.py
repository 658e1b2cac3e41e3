"""CPU: the oracle restatement reproduces the golden vectors generated from the real reference."""
import numpy as np
import pytest
import torch

from helpers import O, build_product, case_inputs, cfg_of, golden, rel_err

CASES = [("tiny_b2_32x64", "trained"), ("tiny_b1_33x47", "trained"), ("tiny_b2_32x64_init", "init"),
         ("w18_b1_32x64", "trained"), ("w48_b1_33x33", "trained"), ("tiny_b2_32x64_nohdz", "trained")]


def _sd(cfg, name, wmode, requires_grad=False):
    g, _ = build_product(cfg)
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    O.fill_state_dict(sd, seed_tag=name, mode=wmode)
    if requires_grad:
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize("name,wmode", CASES)
def test_oracle_matches_reference_golden(name, wmode):
    gold = golden(name)
    cfg = cfg_of(str(gold["cfg"]))
    B, H, W, Z, xt, x2t, x3t, eps_z, code = case_inputs(name, gold)
    sd = _sd(cfg, name, wmode, requires_grad=True)
    losses, x1p, x2p, x3p = O.full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, code)
    got = np.array([float(l) for l in losses])
    np.testing.assert_allclose(got, gold["g_losses"], rtol=2e-5)
    for a, k in ((x1p, "x1p"), (x2p, "x2p"), (x3p, "x3p")):
        assert rel_err(a.detach(), gold[k]) < 2e-5, k
    losses[0].backward()
    names = gold["g_grad_names"].tolist()
    norms = dict(zip(names, gold["g_grad_norms"]))
    worst = 0.0
    for k, v in sd.items():
        if k in norms and v.grad is not None and norms[k] > 1e-8:
            worst = max(worst, abs(float(v.grad.double().norm()) - norms[k]) / norms[k])
    assert worst < 1e-3, worst
    # BN running statistics after the G step
    for k in ("encz_model.bn1.running_mean", "encdec_model.decf_bn2.running_mean", "D_model_frame.bn1.running_var"):
        assert rel_err(sd[k], gold["after:" + k]) < 1e-5, k
    assert int(sd["D_model_frame.bn1.num_batches_tracked"]) == int(gold["after:D_model_frame.bn1.num_batches_tracked"]) == 3
    # D step continues from the updated running stats
    dl = O.full_d_forward({k: v.detach() for k, v in sd.items()}, cfg, x2t, x2p.detach())
    np.testing.assert_allclose(np.array([float(l) for l in dl]), gold["d_losses"], rtol=2e-5)


@pytest.mark.parametrize("name,wmode", CASES[:2])
def test_oracle_eval_prior_sampling(name, wmode):
    gold = golden(name)
    cfg = cfg_of(str(gold["cfg"]))
    B, H, W, Z, xt, x2t, x3t, eps_z, code = case_inputs(name, gold)
    sd = _sd(cfg, name, wmode)
    with torch.no_grad():
        losses, x1p, x2p, x3p = O.full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, code, training=False,
                                                      sampling_mode="prior_sampling")
    np.testing.assert_allclose(np.array([float(l) for l in losses]), gold["eval_losses"], rtol=2e-5)
    assert rel_err(x2p, gold["eval_x2p"]) < 2e-5


def test_oracle_toy_matches_reference_golden():
    import models.toy_fc as T
    import utils.utils as U
    import core.criterion as Cr
    gold = golden("toy_b500")
    cfg = cfg_of("vae2_hrnet_tiny_32x64.yaml")
    g = U.FullToyModel_encdec(T.get_encz_model(cfg), T.get_encdec_model(cfg), T.get_D_model(cfg), Cr.L1Loss(),
                              Cr.KLLoss(), Cr.lsgan_adversarial_loss(), 1.0, 0.1, 1.0, 1.0)
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    O.fill_state_dict(sd, seed_tag="toy_b500", mode="trained")
    xt, x2t, x3t = (torch.from_numpy(gold[k]) for k in ("xt", "x2t", "x3t"))
    eps, code = O.det_normal("toy_b500:eps", (500, 8)), O.det_normal("toy_b500:code", (500, 8))
    losses, x1p, x2p, x3p = O.toy_full_forward(sd, xt, x2t, x3t, eps, code, multiplier=0.5)
    np.testing.assert_allclose(np.array([float(l) for l in losses]), gold["losses"], rtol=2e-5)
    assert rel_err(x3p, gold["x3p"]) < 2e-5


def test_known_answers():
    # KL(mu=0, logvar=0) = 0 ; closed forms (SURVEY.md §8c self-made KATs)
    z = torch.zeros(2, 8, 4, 4)
    assert float(O.kl_loss([z], [z])) == 0.0
    mu, lv = torch.full((2, 1, 1, 1), 2.0), torch.full((2, 1, 1, 1), 1.0)
    assert abs(float(O.kl_loss([mu], [lv])) - 0.5 * (4 + np.e - 1 - 1)) < 1e-6
    assert float(O.l1_loss(torch.ones(4, 3), torch.zeros(4, 3))) == 3.0
    assert float(O.lsgan_loss(torch.zeros(2, 5), "real")) == 5.0
    assert O.branch_sizes(473, 473) == [(473, 473), (237, 237), (119, 119), (60, 60)]


def test_oracle_matches_reference_at_baseline_size_w18_256x512():
    """BASELINE configs[1] size: the oracle's G-step forward (training-mode BN, no autograd: ~15 s) against the
    subsampled fixture the unmodified reference produced (oracle/make_golden.py w18_full)."""
    name = "w18_b1_256x512"
    gold = golden(name)
    cfg = cfg_of(str(gold["cfg"]))
    B, H, W, Z, xt, x2t, x3t, eps_z, code = case_inputs(name, gold)
    sd = _sd(cfg, name, "trained")
    with torch.no_grad():
        losses, x1p, x2p, x3p = O.full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, code)
    np.testing.assert_allclose(np.array([float(l) for l in losses]), gold["g_losses"], rtol=2e-5)
    s = int(gold["sub"])
    for a, k in ((x1p, "x1p"), (x2p, "x2p"), (x3p, "x3p")):
        assert rel_err(a[..., ::s, ::s], gold[k]) < 2e-5, k
        l2 = a.double().pow(2).sum(dim=(0, 2, 3)).sqrt().numpy()
        np.testing.assert_allclose(l2, gold[k + "_l2"], rtol=2e-5)
    for k in ("encz_model.bn1.running_mean", "encdec_model.decf_bn2.running_mean", "D_model_frame.bn1.running_var"):
        assert rel_err(sd[k], gold["after:" + k]) < 1e-5, k
