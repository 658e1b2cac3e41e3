"""CPU oracle for the VAE^2 hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``vae-2_b200/``
does.  It is a *functional restatement* (state_dict in, tensors out) of the
reference's algorithm for the path SURVEY.md §8 names, written against the
CPU kernels of the reference's own arithmetic dependency:

  third-party dependency : PyTorch (un-pinned by the reference's requirements.txt;
                           this image provides torch 2.11.0) -- conv2d, batch_norm,
                           bilinear interpolate (align_corners=False), L1/MSE sums.
  pinned by              : tests/golden/*.npz, produced by oracle/make_golden.py from
                           the UNMODIFIED reference modules (imported from
                           /root/reference in the build container) on procedurally
                           generated weights/inputs with injected eps.  The reference
                           itself ships no golden vectors, so the
                           golden files are "outputs of the reference itself run here".

Every function cites the reference file:line it restates.  All paths are relative
to the reference root.
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.01  # lib/models/enc_hrnet.py:23
BN_EPS = 1e-5       # nn.BatchNorm2d default used at every BN site


# ---------------------------------------------------------------------------
# deterministic, RNG-library-independent tensors (weights, inputs, eps)
# ---------------------------------------------------------------------------
def det_normal(tag, shape, scale=1.0, offset=0.0):
    """Standard normal filled from numpy's legacy MT19937 seeded by crc32(tag);
    bit-stable across numpy/torch versions so both sides can regenerate it."""
    rs = np.random.RandomState(zlib.crc32(tag.encode()) & 0x7FFFFFFF)
    a = rs.standard_normal(size=tuple(shape)).astype(np.float32)
    return torch.from_numpy(a * np.float32(scale) + np.float32(offset))


def det_uniform(tag, shape, lo, hi):
    rs = np.random.RandomState(zlib.crc32(tag.encode()) & 0x7FFFFFFF)
    a = rs.uniform(lo, hi, size=tuple(shape)).astype(np.float32)
    return torch.from_numpy(a)


def fill_state_dict(sd, seed_tag="w", mode="trained"):
    """Overwrite every entry of ``sd`` (a reference-named state_dict) in place.

    mode="trained": trained-scale weights (SURVEY.md §7 'tiny-variance regime'):
      conv ~ N(0, sqrt(2/fan_in)), conv bias ~ N(0,0.1), BN gamma ~ U(0.5,1.5),
      beta ~ N(0,0.1), running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5).
    mode="init": the reference's init (enc_hrnet.py:753-760): conv N(0,1e-3), BN (1,0).
    """
    for k, v in sd.items():
        tag = seed_tag + ":" + k
        if k.endswith("num_batches_tracked"):
            v.zero_()
        elif k.endswith("running_mean"):
            v.copy_(det_normal(tag, v.shape, 0.1) if mode == "trained" else torch.zeros_like(v))
        elif k.endswith("running_var"):
            v.copy_(det_uniform(tag, v.shape, 0.5, 1.5) if mode == "trained" else torch.ones_like(v))
        elif v.dim() == 4 or v.dim() == 2:  # conv / linear weight
            fan_in = int(np.prod(v.shape[1:]))
            std = (2.0 / fan_in) ** 0.5 if mode == "trained" else 1e-3
            if mode == "trained" and "encz_model.last_layer" in k:
                std *= 0.05  # keep logvar = O(0.1) so exp(logvar) in the KL stays well-conditioned
            v.copy_(det_normal(tag, v.shape, std))
        elif k.endswith(".weight"):  # BN gamma (1-D)
            v.copy_(det_uniform(tag, v.shape, 0.5, 1.5) if mode == "trained" else torch.ones_like(v))
        elif k.endswith(".bias"):
            # BN beta or conv/linear bias
            v.copy_(det_normal(tag, v.shape, 0.1) if mode == "trained" else torch.zeros_like(v))
        else:
            raise KeyError("unexpected state_dict entry " + k)
    return sd


def make_clips(tag, B, H, W, L=3):
    """Synthetic (xt, x2t, x3t): N(0,1) frames, temporally correlated."""
    xt = det_normal(tag + ":xt", (B, 3 * L, H, W))
    x2t = xt + 0.1 * det_normal(tag + ":x2t", (B, 3 * L, H, W))
    x3t = x2t + 0.1 * det_normal(tag + ":x3t", (B, 3 * L, H, W))
    return xt, x2t, x3t


def branch_sizes(H, W, n=4):
    """Branch i resolution: stride-2 3x3 p1 conv chain, H_{i+1} = ceil(H_i/2)."""
    out = []
    for _ in range(n):
        out.append((H, W))
        H, W = (H + 1) // 2, (W + 1) // 2
    return out


def make_eps(tag, B, Z, H, W, hd_z=True):
    """eps for the 4 posterior maps (utils.py:92-93 call order) -- or the single [B,Z,1,1] draw without HD_Z
    (utils.py:98-100) -- and the encoder's random code [B,Z,1,1] (enc_hrnet.py:456, drawn after them)."""
    if not hd_z:
        return [det_normal("%s:eps0" % tag, (B, Z, 1, 1))], det_normal(tag + ":code", (B, Z, 1, 1))
    eps_z = [det_normal("%s:eps%d" % (tag, i), (B, Z, h, w)) for i, (h, w) in enumerate(branch_sizes(H, W))]
    code = det_normal(tag + ":code", (B, Z, 1, 1))
    return eps_z, code


# ---------------------------------------------------------------------------
# primitives (the reference's nn.Module call sites, functionally)
# ---------------------------------------------------------------------------
class _Ctx:
    """Carries the state_dict, train/eval flag and an optional activation tap."""

    def __init__(self, sd, training, taps=None):
        self.sd = sd
        self.training = training
        self.taps = taps  # dict name -> tensor, filled when not None

    def tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach().clone()


def _conv(c, name, x, stride=1):
    # nn.Conv2d sites: 3x3 p1 (enc_hrnet.py:27-30) or 1x1 p0; bias only where the module has one
    w = c.sd[name + ".weight"]
    b = c.sd.get(name + ".bias")
    k = w.shape[-1]
    return F.conv2d(x, w, b, stride=stride, padding=k // 2)


def _bn(c, name, x):
    # BatchNorm2d(momentum=0.01): batch stats + running update in train, running stats in eval
    rm, rv = c.sd[name + ".running_mean"], c.sd[name + ".running_var"]
    y = F.batch_norm(x, rm, rv, c.sd[name + ".weight"], c.sd[name + ".bias"],
                     c.training, BN_MOMENTUM, BN_EPS)
    if c.training and (name + ".num_batches_tracked") in c.sd:
        c.sd[name + ".num_batches_tracked"] += 1
    return y


def _basic_block(c, p, x):
    # BasicBlock.forward, enc_hrnet.py:46-62 (never has a downsample in these nets)
    out = F.relu(_bn(c, p + ".bn1", _conv(c, p + ".conv1", x)))
    out = _bn(c, p + ".bn2", _conv(c, p + ".conv2", out))
    return F.relu(out + x)


def _bottleneck(c, p, x):
    # Bottleneck.forward, enc_hrnet.py:83-103
    out = F.relu(_bn(c, p + ".bn1", _conv(c, p + ".conv1", x)))
    out = F.relu(_bn(c, p + ".bn2", _conv(c, p + ".conv2", out)))
    out = _bn(c, p + ".bn3", _conv(c, p + ".conv3", out))
    res = x
    if (p + ".downsample.0.weight") in c.sd:
        res = _bn(c, p + ".downsample.1", _conv(c, p + ".downsample.0", x))
    return F.relu(out + res)


def _block_chain(c, p, x, kind, n):
    blk = _bottleneck if kind == "BOTTLENECK" else _basic_block
    for i in range(n):
        x = blk(c, "%s.%d" % (p, i), x)
    return x


def _hr_module(c, p, xs, scfg):
    # HighResolutionModule.forward, enc_hrnet.py:226-250; fuse layers :177-221
    nb = len(xs)
    xs = [_block_chain(c, "%s.branches.%d" % (p, b), xs[b], scfg["BLOCK"], scfg["NUM_BLOCKS"][b])
          for b in range(nb)]
    if nb == 1:
        return xs
    outs = []
    for i in range(nb):
        acc = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = _bn(c, fp + ".1", _conv(c, fp + ".0", xs[j]))
                t = F.interpolate(t, size=xs[i].shape[-2:], mode="bilinear", align_corners=False)
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _bn(c, "%s.%d.1" % (fp, k), _conv(c, "%s.%d.0" % (fp, k), t, stride=2))
                    if k != i - j - 1:
                        t = F.relu(t)
            acc = t if acc is None else acc + t
        outs.append(F.relu(acc))
    return outs


def _stage(c, p, xs, scfg):
    for m in range(scfg["NUM_MODULES"]):
        xs = _hr_module(c, "%s.%d" % (p, m), xs, scfg)
    return xs


def _transition(c, p, prev, n_cur):
    # _make_transition_layer (enc_hrnet.py:372-406) as used in forward (:796-817):
    # existing branches pass through (or 3x3 s1 conv+BN+ReLU when widths differ),
    # new branches are 3x3 s2 chains fed from the LAST previous branch.
    out = []
    n_pre = len(prev)
    for i in range(n_cur):
        if i < n_pre:
            name = "%s.%d.0" % (p, i)
            if (name + ".weight") in c.sd:
                out.append(F.relu(_bn(c, "%s.%d.1" % (p, i), _conv(c, name, prev[i]))))
            else:
                out.append(prev[i])
        else:
            t = prev[-1]
            for j in range(i + 1 - n_pre):
                t = F.relu(_bn(c, "%s.%d.%d.1" % (p, i, j), _conv(c, "%s.%d.%d.0" % (p, i, j), t, stride=2)))
            out.append(t)
    return out


def _trunk(c, pre, x, extra, cat_maps=None):
    """Stem -> layer1 -> stage2..4 of one HRNet (enc_hrnet.py:787-831 / 849-889 / 1070-1101).

    pre      : '' | 'decf_' | 'decp_'   attribute prefix inside the module
    cat_maps : None, or per-branch list of tensors concatenated IN FRONT of the
               features before ``transition3_e`` (enc_hrnet.py:818-830).
    """
    x = F.relu(_bn(c, pre + "bn1", _conv(c, pre + "conv1", x)))
    x = F.relu(_bn(c, pre + "bn2", _conv(c, pre + "conv2", x)))
    c.tap(pre + "stem", x)
    s1 = extra["STAGE1"]
    x = _block_chain(c, pre + "layer1", x, s1["BLOCK"], s1["NUM_BLOCKS"][0])
    c.tap(pre + "layer1", x)
    xs = _transition(c, pre + "transition1", [x], extra["STAGE2"]["NUM_BRANCHES"])
    ys = _stage(c, pre + "stage2", xs, extra["STAGE2"])
    xs = _transition(c, pre + "transition2", ys, extra["STAGE3"]["NUM_BRANCHES"])
    ys = _stage(c, pre + "stage3", xs, extra["STAGE3"])
    for i, t in enumerate(ys):
        c.tap("%sstage3.%d" % (pre, i), t)
    xs = _transition(c, pre + "transition3", ys, extra["STAGE4"]["NUM_BRANCHES"])
    if cat_maps is not None:
        xs = [torch.cat(list(cat_maps[b]) + [xs[b]], dim=1) for b in range(len(xs))]
        xs = _transition(c, pre + "transition3_e", xs, len(xs))
    ys = _stage(c, pre + "stage4", xs, extra["STAGE4"])
    for i, t in enumerate(ys):
        c.tap("%sstage4.%d" % (pre, i), t)
    return ys


def _upcat(ys):
    # enc_hrnet.py:833-839: bilinear (align_corners=False) to branch-0 size, channel concat
    size = ys[0].shape[-2:]
    return torch.cat([ys[0]] + [F.interpolate(t, size=size, mode="bilinear", align_corners=False)
                                for t in ys[1:]], dim=1)


def _head(c, p, x):
    # last_layer*: 1x1+bias -> BN -> ReLU -> 1x1(+bias)   (enc_hrnet.py:323-338)
    return _conv(c, p + ".3", F.relu(_bn(c, p + ".1", _conv(c, p + ".0", x))))


def _code_maps(code, xs):
    # _gen_code_map, enc_hrnet.py:454-462
    return [code.repeat(1, 1, t.shape[-2], t.shape[-1]) for t in xs]


class bf16_storage:
    """Context manager: emulate bf16 ACTIVATION STORAGE inside this oracle -- every conv and BN output is rounded to
    bf16 (straight-through for autograd), arithmetic stays in the tensors' own dtype (use fp64).  This is the noise
    model of the tensor-core path (bf16 tensors in HBM, fp32 accumulation): tests use it as the yardstick that tells
    bf16 rounding noise, amplified by a deep random-weight net, apart from a wrong kernel."""

    def __enter__(self):
        global _conv, _bn
        self._c, self._b = _conv, _bn
        q = lambda x: x + (x.bfloat16().to(x.dtype) - x).detach()
        oc, ob = _conv, _bn
        _conv = lambda c, name, x, stride=1: q(oc(c, name, x, stride))
        _bn = lambda c, name, x: q(ob(c, name, x))
        return self

    def __exit__(self, *exc):
        global _conv, _bn
        _conv, _bn = self._c, self._b
        return False


# ---------------------------------------------------------------------------
# the four networks
# ---------------------------------------------------------------------------
def encz_forward(sd, cfg, x, training=True, taps=None):
    """HighResolutionNetEDz.forward (HD_Z), enc_hrnet.py:1070-1122 -> 4 maps [B,2Z,H_i,W_i]."""
    c = _Ctx(sd, training, taps)
    ys = _trunk(c, "", x, cfg.MODEL.EXTRA)
    if not cfg.MODEL.EXTRA.HD_Z:
        # non-HD_Z head, enc_hrnet.py:1023-1041 / :1107-1116: avgpool -> 1x1+bias -> BN -> ReLU -> 1x1+bias -> [B,2Z,1,1]
        h = F.adaptive_avg_pool2d(_upcat(ys), (1, 1))
        h = F.relu(_bn(c, "last_layer.2", _conv(c, "last_layer.1", h)))
        return _conv(c, "last_layer.4", h)
    return [_conv(c, "last_layer.%d.0" % i, ys[i]) for i in range(len(ys))]


def dsc_forward(sd, cfg, x, training=True, taps=None):
    """HighResolutionNetDsc (uses base forward, enc_hrnet.py:464-507, head :1136-1151)."""
    c = _Ctx(sd, training, taps)
    ys = _trunk(c, "", x, cfg.MODEL.EXTRA)
    return _head(c, "last_layer", _upcat(ys))


def _one_net(c, pre, x, extra, z, code, is_encoder):
    """_encoder_foward / _decoder_*_foward, enc_hrnet.py:787-963 (HD_Z, not baseline)."""
    nb = extra["STAGE4"]["NUM_BRANCHES"]
    H, W = x.shape[-2:]
    sizes = branch_sizes(H, W, nb)
    if torch.is_tensor(z):      # no HD_Z: one per-sample z [B,Z,1,1] repeated over every map (_gen_code_map(x_list, z), :819)
        z = [z.repeat(1, 1, h, w) for (h, w) in sizes]
    if is_encoder:
        cmaps = [[code.repeat(1, 1, h, w), z[b]] for b, (h, w) in enumerate(sizes)]
    else:
        cmaps = [[z[b]] for b in range(nb)]
    ys = _trunk(c, pre, x, extra, cat_maps=cmaps)
    feat = _upcat(ys)
    return torch.cat([_head(c, "%slast_layer_%d" % (pre, i), feat) for i in (1, 2, 3)], dim=1)


def encdec_forward(sd, cfg, x, z, code, training=True, taps=None):
    """HighResolutionNetED.forward, enc_hrnet.py:965-981 -> (x1_pred, x2_pred, x3_pred).
    Order matters for BN running stats: encoder, decoder-future, decoder-past."""
    c = _Ctx(sd, training, taps)
    extra = cfg.MODEL.EXTRA
    x2p = _one_net(c, "", x, extra, z, code, True)
    x3p = _one_net(c, "decf_", x2p, extra, z, None, False)
    x1p = _one_net(c, "decp_", x2p, extra, z, None, False)
    return x1p, x2p, x3p


# ---------------------------------------------------------------------------
# loss primitives and the ELBO assembly
# ---------------------------------------------------------------------------
def l1_loss(pred, target):
    # L1Loss.forward, lib/core/criterion.py:66-69
    return (pred - target).abs().sum() / pred.shape[0]


def kl_loss(mus, logvars):
    # KLLoss.forward (list form), lib/core/criterion.py:76-87
    tot = 0.0
    for m, v in zip(mus, logvars):
        tot = tot + torch.sum(0.5 * (m ** 2 + torch.exp(v) - v - 1)) / m.shape[0]
    return tot


def lsgan_loss(sample, mode):
    # lsgan_adversarial_loss.forward, lib/core/criterion.py:96-103
    tgt = 1.0 if mode == "real" else 0.0
    return ((sample - tgt) ** 2).sum() / sample.shape[0]


def reparam(mus, logvars, eps):
    # z = mu + exp(0.5*logvar) * eps, lib/utils/utils.py:92-93
    return [m + torch.exp(0.5 * v) * e for m, v, e in zip(mus, logvars, eps)]


def split_sd(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, code, multiplier=1.0,
                        baseline_mode="VAE_NATIVE", sampling_mode="default",
                        training=True, with_gan=True, lambdas=None, taps=None):
    """FullModel_encdec.forward (not baseline), lib/utils/utils.py:67-155.

    ``sd`` is the wrapper's state_dict (prefixes encz_model./encdec_model./
    D_model_sequence./D_model_frame.).  Returns (losses[7], x1p, x2p, x3p).
    with_gan=False gives the ELBO-only sub-step (gan terms 0).
    """
    l1w, l2w, l3w, lg = lambdas or (cfg.TRAIN.X1RECON_LAMBDA, cfg.TRAIN.X2RECON_LAMBDA,
                                    cfg.TRAIN.X3RECON_LAMBDA, cfg.TRAIN.GAN_LAMBDA)
    Z = cfg.MODEL.EXTRA.Z_DIM
    L = cfg.TRAIN.CLIP_LENGTH
    klw = l3w * multiplier if baseline_mode == "VAE_ANNEAL" else l3w  # utils.py:74
    muvars = encz_forward(split_sd(sd, "encz_model."), cfg, torch.cat([xt, x3t], 1), training, taps)
    hd = cfg.MODEL.EXTRA.HD_Z
    if not hd:                   # utils.py:82-83, 96-100: one [B,2Z,1,1] tensor, eps_z one [B,Z,1,1] tensor
        muvars, eps_z = [muvars], ([eps_z] if torch.is_tensor(eps_z) else list(eps_z))
    mus = [mv[:, :Z] for mv in muvars]
    logvars = [mv[:, Z:] for mv in muvars]
    if sampling_mode == "prior_sampling":
        z = list(eps_z)  # utils.py:88-90
    else:
        z = reparam(mus, logvars, eps_z)
    if not hd:
        z = z[0]
    x1p, x2p, x3p = encdec_forward(split_sd(sd, "encdec_model."), cfg, xt, z, code, training, taps)
    lx1, lx2, lx3 = l1_loss(x1p, xt), l1_loss(x2p, x2t), l1_loss(x3p, x3t)
    kl = kl_loss(mus, logvars)
    if with_gan:
        gseq = 0.5 * lsgan_loss(dsc_forward(split_sd(sd, "D_model_sequence."), cfg, x2p, training), "real")
        gfrm = 0.0
        dsd = split_sd(sd, "D_model_frame.")
        for f in range(x2t.shape[1] // L):  # utils.py:116-119
            gfrm = gfrm + 0.5 * lsgan_loss(dsc_forward(dsd, cfg, x2p[:, 3 * f:3 * f + 3], training), "real")
    else:
        gseq = torch.zeros(())
        gfrm = torch.zeros(())
    total = l1w * lx1 + l2w * lx2 + l3w * lx3 + klw * kl + lg * (gseq + gfrm)  # utils.py:150-152
    return [total.reshape(1), lx1, lx2, lx3, kl, gseq, gfrm], x1p, x2p, x3p


def full_d_forward(sd, cfg, x2t, x2p, training=True, gan_lambda=1.0):
    """FullModel_D.forward, lib/utils/utils.py:259-276 -> [D_losses[1], D_seq, D_frm].
    Call order (matters for BN running stats): seq(real), seq(fake), then per frame real, fake."""
    L = cfg.TRAIN.CLIP_LENGTH
    ssd, fsd = split_sd(sd, "D_model_sequence."), split_sd(sd, "D_model_frame.")
    x2t, x2p = x2t.detach(), x2p.detach()
    rs = 0.5 * lsgan_loss(dsc_forward(ssd, cfg, x2t, training), "real")
    fs = 0.5 * lsgan_loss(dsc_forward(ssd, cfg, x2p, training), "fake")
    rf = ff = 0.0
    for f in range(x2t.shape[1] // L):
        rf = rf + 0.5 * lsgan_loss(dsc_forward(fsd, cfg, x2t[:, 3 * f:3 * f + 3], training), "real")
        ff = ff + 0.5 * lsgan_loss(dsc_forward(fsd, cfg, x2p[:, 3 * f:3 * f + 3], training), "fake")
    dseq, dfrm = rs + fs, rf + ff
    return [(gan_lambda * (dseq + dfrm)).reshape(1), dseq, dfrm]


# ---------------------------------------------------------------------------
# toy VAE^2 (BASELINE config #1): lib/models/toy_fc.py + FullToyModel_encdec
# ---------------------------------------------------------------------------
def _mlp(sd, p, x):
    # toy_fc.forward: h1 -> h2 -> output (toy_fc.py:58-61)
    h = F.relu(F.linear(x, sd[p + "h1.0.weight"], sd[p + "h1.0.bias"]))
    h = F.relu(F.linear(h, sd[p + "h2.0.weight"], sd[p + "h2.0.bias"]))
    return F.linear(h, sd[p + "output.weight"], sd[p + "output.bias"])


def toy_full_forward(sd, xt, x2t, x3t, eps, code, multiplier=1.0, lambdas=(1.0, 0.1, 1.0, 1.0), Z=8):
    """FullToyModel_encdec.forward (not baseline), lib/utils/utils.py:185-241."""
    l1w, l2w, l3w, lg = lambdas
    mv = _mlp(split_sd(sd, "encz_model."), "", torch.cat([xt, x3t], 1))
    mu, lv = mv[:, :Z], mv[:, Z:]
    z = mu + torch.exp(0.5 * lv) * eps
    ed = split_sd(sd, "encdec_model.")
    x2p = _mlp(ed, "", torch.cat([xt, z, code], -1))      # toy_fc.py:108-119 (z first, random code second)
    x1p = _mlp(ed, "decp_", torch.cat([x2p, z], -1))       # toy_fc.py:130-137
    x3p = _mlp(ed, "decf_", torch.cat([x2p, z], -1))       # toy_fc.py:121-128
    lx1, lx2, lx3 = l1_loss(x1p, xt), l1_loss(x2p, x2t), l1_loss(x3p, x3t)
    kl = torch.sum(0.5 * (mu ** 2 + torch.exp(lv) - lv - 1)) / mu.shape[0]
    gan = lsgan_loss(_mlp(split_sd(sd, "D_model."), "", x2p), "real")
    total = l1w * lx1 + (l2w * multiplier) * lx2 + l3w * lx3 + l3w * kl + lg * gan  # utils.py:193,235-237
    return [total.reshape(1), lx1, lx2, lx3, kl, gan, gan], x1p, x2p, x3p
