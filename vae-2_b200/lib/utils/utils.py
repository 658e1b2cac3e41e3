"""Loss-in-forward wrappers of VAE^2 on the B200 engine  --  drop-in for the reference's
lib/utils/utils.py (``FullModel_encdec`` :39-155, ``FullToyModel_encdec`` :158-241,
``FullModel_D`` :244-276, ``FullToyModel_D`` :279-299 and the host helpers :355-468).

Constructor arguments, forward keywords and the returned
``([loss_all[1], x1_recon, x2_recon, x3_recon, KL, gan_seq, gan_frm], x1p, x2p, x3p)`` are the
reference's.  Internally a G-step is: posterior net -> ONE launch for the four
reparameterisations + KL -> encoder/decoders -> discriminators -> ONE launch for the three L1
terms and the four LSGAN terms; the reference's 14 isnan/isinf host syncs per step
(:63-65, :94, :107) become device counters checked lazily (engine.check_finite).
"""
import logging
import math
import os
import sys
import time
from pathlib import Path

import torch
import torch.nn as nn

_LIB = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _LIB not in sys.path:
    sys.path.insert(0, _LIB)
from _engine_loader import engine  # noqa: E402

import _fallthrough  # noqa: E402

_E = engine()


def __getattr__(name):
    """PEP 562: names this module does not define (FullModel, FullModel_all, get_confusion_matrix -- the legacy
    segmentation path, reference :21-37, :302-352, :434-457) come from the reference's own utils.utils when its lib/
    is on sys.path behind this tree."""
    if name.startswith("__"):
        raise AttributeError(name)
    return _fallthrough.reference_attr("utils.utils", name)


def _stack_D():
    """Run the discriminator calls of a step as stacked passes (models.enc_hrnet.HighResolutionNetDsc.forward_groups);
    VAE2_STACK_D=0 restores one pass per call."""
    return os.environ.get("VAE2_STACK_D", "1") != "0"


class _frozen:
    """Parameters of `modules` do not require grad inside the block (the plan recorder then emits no weight /
    gamma / beta gradient for them)."""

    def __init__(self, modules, on):
        self.ps = [p for m in modules for p in m.parameters() if p.requires_grad] if on else []

    def __enter__(self):
        for p in self.ps:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.ps:
            p.requires_grad_(True)
        return False


def _eager_gan():
    """Evaluate the GAN terms eagerly (forward AND backward of every discriminator pass inside the wrapper's forward):
    a discriminator's activations then live only for the duration of its own pass (engine scratch region) instead of
    until the caller's backward().  VAE2_EAGER_GAN=0 restores plain autograd ordering."""
    return os.environ.get("VAE2_EAGER_GAN", "1") != "0"


def _d_stack():
    """Most discriminator calls stacked into one pass (memory of a stacked pass grows with the count)."""
    return max(1, int(os.environ.get("VAE2_D_STACK", "6")))


def _eager_pass(net, srcs, targets, B, phase, grad_wrt):
    """One stacked discriminator pass evaluated forward AND backward: srcs = [(tensor, channel offset)], targets = LSGAN
    targets per call.  Returns (sum of the 0.5/B-scaled terms, gradients w.r.t. ``grad_wrt``)."""
    with _E.activation_phase("scratch:" + phase):
        out = net.forward_groups(srcs) if len(srcs) > 1 or srcs[0][1] != 0 else net(srcs[0][0])
        spec = [dict(kind=2, slot=0, a=i, b=None, scale=0.5 / B, target=t, name="%s%d" % (phase, i)) for i, t in enumerate(targets)]
        v = _E.elbo_terms(spec, 1, [out[i * B:(i + 1) * B] for i in range(len(srcs))])[0][0]
        g = torch.autograd.grad(v, grad_wrt, allow_unused=True)
    return v.detach(), g


def _add(a, b):
    return b if a is None else (a if b is None else a + b)


class _EagerGanG(torch.autograd.Function):
    """GAN terms of the generator step (reference lib/utils/utils.py:114-119) with the discriminators' parameters
    frozen: d(term)/d(x2t_predict) is computed right away -- the terms are linear in the upstream gradient, so the
    caller's backward() only scales the two stored gradients.  At most VAE2_D_STACK frames share a stacked pass."""

    @staticmethod
    def forward(ctx, wrapper, x2p):
        B = x2p.shape[0]
        L = wrapper.D_model_sequence.clip_length
        nf = x2p.shape[1] // L
        with torch.enable_grad(), _frozen([wrapper.D_model_sequence, wrapper.D_model_frame], True):
            xd = x2p.detach().requires_grad_(True)
            v_seq, (g_seq,) = _eager_pass(wrapper.D_model_sequence, [(xd, 0)], [1.0], B, "Gseq", [xd])
            per, v_frm, g_frm = min(nf, _d_stack()), None, None
            for f0 in range(0, nf, per):
                fr = list(range(f0, min(nf, f0 + per)))
                if per == nf:
                    srcs = [(xd, 3 * f) for f in fr]                  # one plan reads all windows of x2t_predict
                else:                                                 # chunks replay ONE plan: windows cut here
                    srcs = [(xd[:, 3 * f:3 * f + 3].contiguous(), 0) for f in fr]
                v, (g,) = _eager_pass(wrapper.D_model_frame, srcs, [1.0] * len(fr), B, "Gfrm", [xd])
                v_frm, g_frm = _add(v_frm, v), _add(g_frm, g)
        ctx.save_for_backward(g_seq, g_frm)
        return v_seq, v_frm

    @staticmethod
    def backward(ctx, up_seq, up_frm):
        a, b = ctx.saved_tensors
        out = None
        for up, t in ((up_seq, a), (up_frm, b)):
            if up is not None:
                out = _add(out, up * t)
        return None, out


class _EagerGanD(torch.autograd.Function):
    """Discriminator step (reference lib/utils/utils.py:259-276): every (stacked) pass runs forward, its LSGAN terms
    and backward at once; parameter gradients for a unit upstream gradient are kept (a few MB) and scaled in backward().
    Calls are stacked in the reference's order (seq: real, fake; frames: real_f, fake_f, ...), at most VAE2_D_STACK per
    pass, so BN running statistics see the same update sequence."""

    @staticmethod
    def forward(ctx, wrapper, real, fake, n_seq, *params):
        B = real.shape[0]
        L = wrapper.D_model_sequence.clip_length
        nf = real.shape[1] // L
        seq_p = [p for p in params[:n_seq] if p.requires_grad]
        frm_p = [p for p in params[n_seq:] if p.requires_grad]
        stack = _d_stack()
        with torch.enable_grad():
            calls = [(real, 1.0), (fake, 0.0)]
            v_seq, g_seq = None, [None] * len(seq_p)
            for c0 in range(0, 2, min(2, stack)):
                ch = calls[c0:c0 + min(2, stack)]
                v, g = _eager_pass(wrapper.D_model_sequence, [(t, 0) for t, _ in ch], [tg for _, tg in ch], B, "Dseq", seq_p)
                v_seq, g_seq = _add(v_seq, v), [_add(a, b) for a, b in zip(g_seq, g)]
            # frame calls: windows are cut here (offset 0 in the plan) so that every chunk replays ONE plan
            calls = [(t[:, 3 * f:3 * f + 3].contiguous(), tg) for f in range(nf) for t, tg in ((real, 1.0), (fake, 0.0))]
            v_frm, g_frm = None, [None] * len(frm_p)
            for c0 in range(0, len(calls), stack):
                ch = calls[c0:c0 + stack]
                v, g = _eager_pass(wrapper.D_model_frame, [(t, 0) for t, _ in ch], [tg for _, tg in ch], B, "Dfrm", frm_p)
                v_frm, g_frm = _add(v_frm, v), [_add(a, b) for a, b in zip(g_frm, g)]
        ctx.n_seq, ctx.req = n_seq, [p.requires_grad for p in params]
        ctx.grads = list(g_seq) + list(g_frm)
        return v_seq, v_frm

    @staticmethod
    def backward(ctx, up_seq, up_frm):
        out, it = [], iter(ctx.grads)
        for i, need in enumerate(ctx.req):
            if not need:
                out.append(None)
                continue
            g = next(it)
            up = up_seq if i < ctx.n_seq else up_frm
            out.append(None if (g is None or up is None) else g * up)
        ctx.grads = None
        return (None, None, None, None) + tuple(out)


def _crit():
    from core import criterion as C_
    return C_


def _fast_losses(*crits):
    """True when the criteria are this package's fused ones (otherwise call them as given)."""
    from core import criterion as C_
    kinds = (C_.L1Loss, C_.KLLoss, C_.lsgan_adversarial_loss)
    return all(isinstance(c, k) for c, k in zip(crits, kinds))


class FullModel_encdec(nn.Module):
    def __init__(self, encz_model, encdec_model, D_model_sequence, D_model_frame,
                 criterion_recon, criterion_KL, criterion_gan,
                 x1recon_lambda=1.0, x2recon_lambda=1.0, x3recon_lambda=1.0, gan_lambda=1.0):
        super().__init__()
        self.encz_model, self.encdec_model = encz_model, encdec_model
        self.D_model_sequence, self.D_model_frame = D_model_sequence, D_model_frame
        self.criterion_recon, self.criterion_KL, self.criterion_gan = criterion_recon, criterion_KL, criterion_gan
        self.x1recon_lambda, self.x2recon_lambda = x1recon_lambda, x2recon_lambda
        self.x3recon_lambda, self.gan_lambda = x3recon_lambda, gan_lambda
        # The reference's loop zeroes the discriminators' gradients right after this wrapper's backward
        # (lib/core/function.py:499-512: optimizer_D.zero_grad() precedes loss_D.backward()), so the D weight / gamma /
        # beta gradients of the generator step are dead.  True = do not compute them (their .grad stays as it was).
        self.skip_dead_D_grads = os.environ.get("VAE2_SKIP_DEAD_D_GRADS", "0") == "1"

    def _anomoly_detection(self, tensor_dict=None):
        """Reference semantics (:63-65) without the per-tensor host sync: raises for any earlier
        launch whose device-side counter saw inf/nan."""
        _E.check_finite()

    def forward(self, xt, x2t, x3t, multiplier, is_baseline=False, baseline_mode="VAE_NATIVE",
                sampling_mode="default", xt_last=None, x3t_last=None, eps=None):
        # generator phase: its networks share activation memory with the discriminator phase (engine.ActArena)
        with _E.activation_phase("G"):
            return self._forward(xt, x2t, x3t, multiplier, is_baseline, baseline_mode, sampling_mode, xt_last, x3t_last,
                                 eps)

    def _forward(self, xt, x2t, x3t, multiplier, is_baseline=False, baseline_mode="VAE_NATIVE",
                 sampling_mode="default", xt_last=None, x3t_last=None, eps=None):
        assert sampling_mode in ["default", "prior_sampling", "momentum_sampling"]
        if sampling_mode == "momentum_sampling":
            assert xt_last is not None
            assert x3t_last is not None
        if is_baseline or baseline_mode == "DETERMINISTIC":
            raise NotImplementedError("vae2_b200: the baseline ablations are outside the built hot path")
        baseline_mode = baseline_mode or "VAE_NATIVE"
        self._anomoly_detection()
        B, Z = xt.shape[0], self.encz_model.z_dim
        kl_w = self.x3recon_lambda * multiplier if baseline_mode == "VAE_ANNEAL" else self.x3recon_lambda
        prior = sampling_mode == "prior_sampling"

        muvars = self.encz_model(x=torch.cat([xt, x3t], 1))                       # reference :77
        hd = self.encz_model.hd_z
        if not hd:                                 # one [B, 2Z, 1, 1] tensor (reference :82-83, :96-100)
            muvars = [muvars]
            if eps is not None and torch.is_tensor(eps):
                eps = [eps]
        # eps in the reference's draw order: one randn per posterior map (:89-93)
        if eps is None:
            eps = [torch.randn(mv.shape[0], Z, mv.shape[2], mv.shape[3], device=mv.device) for mv in muvars]
        tensors = list(muvars) + list(eps)
        n = len(muvars)
        spec = [dict(kind=1, slot=0, a=n + i, b=i, scale=1.0 / B, want_z=True, prior=prior, name=i)
                for i in range(n)]
        out = _E.elbo_terms(spec, 1, tensors)
        z_KL_loss, z = out[0][0], list(out[1:])
        if not hd:
            z = z[0]

        xt_predict, x2t_predict, x3t_predict = self.encdec_model(x=xt, z=z, is_baseline=False)   # :105

        L = self.D_model_sequence.clip_length
        nf = x2t.shape[1] // L
        fast = _fast_losses(self.criterion_recon, self.criterion_KL, self.criterion_gan)
        l1_spec = [dict(kind=0, slot=0, a=0, b=1, scale=1.0 / B, name="xt_predict"),
                   dict(kind=0, slot=1, a=2, b=3, scale=1.0 / B, name="x2t_predict"),
                   dict(kind=0, slot=2, a=4, b=5, scale=1.0 / B, name="x3t_predict")]
        l1_in = [xt_predict, xt.detach(), x2t_predict, x2t.detach(), x3t_predict, x3t.detach()]
        eager = (fast and self.skip_dead_D_grads and _eager_gan() and _stack_D() and torch.is_grad_enabled()
                 and x2t_predict.requires_grad and hasattr(self.D_model_frame, "forward_groups"))
        if eager:
            # discriminators frozen (their gradients are dead, see __init__) and evaluated forward+backward right here
            x2t_gan_sequence_loss, x2t_gan_frame_loss = _EagerGanG.apply(self, x2t_predict)
            vals = _E.elbo_terms(l1_spec, 3, l1_in)[0]
            xt_recon_loss, x2t_recon_loss, x3t_recon_loss = vals[0], vals[1], vals[2]
        else:
            with _frozen([self.D_model_sequence, self.D_model_frame], self.skip_dead_D_grads and torch.is_grad_enabled()):
                d_seq = self.D_model_sequence(x2t_predict)
                if _stack_D() and hasattr(self.D_model_frame, "forward_groups"):
                    allf = self.D_model_frame.forward_groups([(x2t_predict, 3 * f) for f in range(nf)])     # :116-119
                    d_frm = [allf[f * B:(f + 1) * B] for f in range(nf)]
                else:
                    d_frm = [self.D_model_frame(x2t_predict[:, f * 3: f * 3 + 3, :, :]) for f in range(nf)]
            if fast:
                spec = l1_spec + [dict(kind=2, slot=3, a=6, b=None, scale=0.5 / B, target=1.0, name="d_seq")]
                spec += [dict(kind=2, slot=4, a=7 + f, b=None, scale=0.5 / B, target=1.0, name="d_frm%d" % f)
                         for f in range(len(d_frm))]
                vals = _E.elbo_terms(spec, 5, l1_in + [d_seq] + d_frm)[0]
                xt_recon_loss, x2t_recon_loss, x3t_recon_loss = vals[0], vals[1], vals[2]
                x2t_gan_sequence_loss, x2t_gan_frame_loss = vals[3], vals[4]
            else:
                xt_recon_loss = self.criterion_recon(predict=xt_predict, target=xt)
                x2t_recon_loss = self.criterion_recon(predict=x2t_predict, target=x2t)
                x3t_recon_loss = self.criterion_recon(predict=x3t_predict, target=x3t)
                x2t_gan_sequence_loss = 0.5 * self.criterion_gan(sample=d_seq, mode="real")
                x2t_gan_frame_loss = torch.sum(torch.stack(
                    [0.5 * self.criterion_gan(sample=d, mode="real") for d in d_frm], 0), 0)

        losses_all = self.x1recon_lambda * xt_recon_loss + self.x2recon_lambda * x2t_recon_loss + \
            self.x3recon_lambda * x3t_recon_loss + kl_w * z_KL_loss + \
            self.gan_lambda * (x2t_gan_sequence_loss + x2t_gan_frame_loss)                       # :150-152
        if _E.finite_check_mode() == "eager":
            # reference :94, :107 assert BEFORE backward / optimizer.step; one blocking read of this step's
            # device-side counters (z maps and the three predictions) keeps that guarantee
            _E.check_finite(block=True)
        return [torch.unsqueeze(losses_all, 0), xt_recon_loss, x2t_recon_loss, x3t_recon_loss, z_KL_loss,
                x2t_gan_sequence_loss, x2t_gan_frame_loss], xt_predict, x2t_predict, x3t_predict


class FullModel_D(nn.Module):
    def __init__(self, D_model_sequence, D_model_frame, criterion_gan, gan_lambda=1.0):
        super().__init__()
        self.D_model_sequence, self.D_model_frame = D_model_sequence, D_model_frame
        self.criterion_gan, self.gan_lambda = criterion_gan, gan_lambda

    def forward(self, x2t, x2t_predict):
        with _E.activation_phase("D"):
            return self._forward(x2t, x2t_predict)

    def _forward(self, x2t, x2t_predict):
        real, fake = x2t.detach(), x2t_predict.detach()
        B = real.shape[0]
        L = self.D_model_sequence.clip_length
        # call order as the reference (:260-267): seq(real), seq(fake), then per frame real, fake
        nf = x2t.shape[1] // L
        if (_eager_gan() and _stack_D() and torch.is_grad_enabled() and hasattr(self.D_model_frame, "forward_groups")
                and isinstance(self.criterion_gan, _crit().lsgan_adversarial_loss)):
            ps, pf = list(self.D_model_sequence.parameters()), list(self.D_model_frame.parameters())
            D_losses_sequence, D_losses_frame = _EagerGanD.apply(self, real, fake, len(ps), *ps, *pf)
            D_losses = self.gan_lambda * (D_losses_sequence + D_losses_frame)
            return [torch.unsqueeze(D_losses, 0), D_losses_sequence, D_losses_frame]
        if _stack_D() and hasattr(self.D_model_frame, "forward_groups"):
            seq = self.D_model_sequence.forward_groups([(real, 0), (fake, 0)])
            frm = self.D_model_frame.forward_groups([(t, 3 * f) for f in range(nf) for t in (real, fake)])
            outs = [seq[:B], seq[B:]] + [frm[i * B:(i + 1) * B] for i in range(2 * nf)]
        else:
            outs = [self.D_model_sequence(real), self.D_model_sequence(fake)]
            for f in range(nf):
                outs.append(self.D_model_frame(real[:, f * 3: f * 3 + 3, :, :]))
                outs.append(self.D_model_frame(fake[:, f * 3: f * 3 + 3, :, :]))
        spec = [dict(kind=2, slot=0, a=0, b=None, scale=0.5 / B, target=1.0, name="d_seq_real"),
                dict(kind=2, slot=0, a=1, b=None, scale=0.5 / B, target=0.0, name="d_seq_fake")]
        for f in range(nf):
            spec.append(dict(kind=2, slot=1, a=2 + 2 * f, b=None, scale=0.5 / B, target=1.0, name="d_frm_real"))
            spec.append(dict(kind=2, slot=1, a=3 + 2 * f, b=None, scale=0.5 / B, target=0.0, name="d_frm_fake"))
        vals = _E.elbo_terms(spec, 2, outs)[0]
        D_losses_sequence, D_losses_frame = vals[0], vals[1]
        D_losses = self.gan_lambda * (D_losses_sequence + D_losses_frame)
        return [torch.unsqueeze(D_losses, 0), D_losses_sequence, D_losses_frame]


class FullToyModel_encdec(nn.Module):
    def __init__(self, encz_model, encdec_model, D_model, criterion_recon, criterion_KL, criterion_gan,
                 x1recon_lambda=1.0, x2recon_lambda=1.0, x3recon_lambda=1.0, gan_lambda=1.0):
        super().__init__()
        self.encz_model, self.encdec_model, self.D_model = encz_model, encdec_model, D_model
        self.criterion_recon, self.criterion_KL, self.criterion_gan = criterion_recon, criterion_KL, criterion_gan
        self.x1recon_lambda, self.x2recon_lambda = x1recon_lambda, x2recon_lambda
        self.x3recon_lambda, self.gan_lambda = x3recon_lambda, gan_lambda

    def forward(self, xt, x2t, x3t, multiplier, is_baseline=False, baseline_mode=None, sampling_mode="default",
                xt_last=None, x3t_last=None, eps=None):
        assert sampling_mode in ["default", "prior_sampling", "momentum_sampling"]
        if is_baseline or baseline_mode == "DETERMINISTIC":
            raise NotImplementedError("vae2_b200: baseline ablations are outside the built hot path")
        _E.check_finite()
        B, Z = xt.shape[0], self.encz_model.z_dim
        x2w = self.x2recon_lambda * multiplier                                     # reference :193
        src = torch.cat([xt_last, x3t_last], 1) if sampling_mode == "momentum_sampling" else torch.cat([xt, x3t], 1)
        muvars = self.encz_model(x=src)                                            # [B, 2Z]
        if eps is None:
            eps = torch.randn(B, Z, device=xt.device)
        out = _E.elbo_terms([dict(kind=1, slot=0, a=1, b=0, scale=1.0 / B, want_z=True,
                                  prior=sampling_mode == "prior_sampling", name="0")], 1,
                            [muvars.reshape(B, 2 * Z, 1, 1), eps.reshape(B, Z, 1, 1)])
        kl, z = out[0][0], out[1].reshape(B, Z)
        x1p, x2p, x3p = self.encdec_model(x=xt, z=z)
        d = self.D_model(x2p)
        tensors = [x1p, xt, x2p, x2t, x3p, x3t, d]
        spec = [dict(kind=0, slot=0, a=0, b=1, scale=1.0 / B, name="xt_predict"),
                dict(kind=0, slot=1, a=2, b=3, scale=1.0 / B, name="x2t_predict"),
                dict(kind=0, slot=2, a=4, b=5, scale=1.0 / B, name="x3t_predict"),
                dict(kind=2, slot=3, a=6, b=None, scale=1.0 / B, target=1.0, name="d")]
        v = _E.elbo_terms(spec, 4, [t if i % 2 == 0 else t.detach() for i, t in enumerate(tensors)])[0]
        total = self.x1recon_lambda * v[0] + x2w * v[1] + self.x3recon_lambda * v[2] + \
            self.x3recon_lambda * kl + self.gan_lambda * v[3]                      # reference :235-237
        return [torch.unsqueeze(total, 0), v[0], v[1], v[2], kl, v[3], v[3]], x1p, x2p, x3p


class FullToyModel_D(nn.Module):
    def __init__(self, D_model, criterion_gan, gan_lambda=1.0):
        super().__init__()
        self.D_model, self.criterion_gan, self.gan_lambda = D_model, criterion_gan, gan_lambda

    def forward(self, x2t, x2t_predict):
        B = x2t.shape[0]
        outs = [self.D_model(x2t.detach()), self.D_model(x2t_predict.detach())]
        spec = [dict(kind=2, slot=0, a=0, b=None, scale=0.5 / B, target=1.0, name="real"),
                dict(kind=2, slot=0, a=1, b=None, scale=0.5 / B, target=0.0, name="fake")]
        d = _E.elbo_terms(spec, 1, outs)[0][0]
        return [torch.unsqueeze(d, 0), d, d]


# ---- host-side helpers the reference's drivers import from this module (:355-468) -------------
def get_world_size():
    return torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1


def get_rank():
    return torch.distributed.get_rank() if torch.distributed.is_initialized() else 0


class AverageMeter(object):
    """Running weighted mean (reference :365-396)."""

    def __init__(self):
        self.initialized, self.val, self.avg, self.sum, self.count = False, None, None, None, None

    def update(self, val, weight=1):
        if not self.initialized:
            self.val, self.avg, self.sum, self.count, self.initialized = val, val, val * weight, weight, True
        else:
            self.val = val
            self.sum += val * weight
            self.count += weight
            self.avg = self.sum / self.count

    def value(self):
        return self.val

    def average(self):
        return self.avg


def create_logger(cfg, cfg_name, phase="train"):
    """Output/log/tensorboard directories and a root logger (reference :398-432)."""
    out_root = Path(cfg.OUTPUT_DIR)
    out_root.mkdir(parents=True, exist_ok=True)
    stem = os.path.basename(cfg_name).split(".")[0]
    final_dir = out_root / cfg.DATASET.DATASET / stem
    final_dir.mkdir(parents=True, exist_ok=True)
    stamp = time.strftime("%Y-%m-%d-%H-%M")
    logging.basicConfig(filename=str(final_dir / "{}_{}_{}.log".format(stem, stamp, phase)),
                        format="%(asctime)-15s %(message)s")
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    logging.getLogger("").addHandler(logging.StreamHandler())
    tb_dir = Path(cfg.LOG_DIR) / cfg.DATASET.DATASET / cfg.MODEL.NAME / (stem + "_" + stamp)
    tb_dir.mkdir(parents=True, exist_ok=True)
    return logger, str(final_dir), str(tb_dir)


def adjust_learning_rate(optimizer, base_lr, max_iters, cur_iters, power=0.9):
    lr = base_lr * ((1 - float(cur_iters) / max_iters) ** power)
    optimizer.param_groups[0]["lr"] = lr
    return lr


def dynamic_coeff(max_iters, cur_iters):
    return math.sin((math.pi / 2) * (float(cur_iters) / float(max_iters)))
