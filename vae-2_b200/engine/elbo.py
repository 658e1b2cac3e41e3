"""Autograd front-end of the fused ELBO-terms kernel (csrc/elbo.cu).

One launch evaluates a table of terms -- L1 reconstruction, reparameterisation + KL,
LSGAN -- on the reference's boundary tensors (NCHW fp32) and returns the per-slot sums as a
device tensor; no host synchronisation (the reference's isnan/isinf asserts,
lib/utils/utils.py:63-65, become device counters read lazily, see ``check_finite``).
"""
import ctypes as C

import torch

from . import native as N

import os

_pending_flags = []   # (names, device int32 tensor, event) of launches whose finite-flags were not read yet
MAX_PENDING = 16      # callers that never check (FullModel_D, stand-alone criteria) are drained beyond this
LAUNCHES = {"n": 0}


def finite_check_mode():
    """'eager' (default): the wrappers read this step's flags BEFORE returning the losses, so a nan/inf raises
    before backward / optimizer.step like the reference's _anomoly_detection (lib/utils/utils.py:63-65), at the
    cost of ONE host sync per step instead of the reference's 14.  'lazy': flags are read at the next wrapper
    call (no sync at all; a bad step is reported one step late)."""
    return os.environ.get("VAE2_FINITE_CHECK", "eager")


def _table(structs, dev):
    raw = bytearray(bytes(structs))
    return torch.frombuffer(raw, dtype=torch.uint8).to(dev)


def check_finite(block=False):
    """Raise AssertionError (as the reference's _anomoly_detection does) if an earlier launch saw
    inf/nan in z or in a prediction.  Flags of launches that have already completed cost no stall."""
    keep, failed = [], None
    for names, flags, ev in _pending_flags:
        if block or ev.query():
            for n, b in zip(names, flags.tolist()):
                if b != 0 and failed is None:
                    failed = n
        else:
            keep.append((names, flags, ev))
    _pending_flags[:] = keep        # a reported launch is consumed even when it raises
    assert failed is None, "{} got nan or inf".format(failed)
    if block:
        from . import peer
        peer.check()                # multi-GPU: a SyncBN launch that gave up waiting for a peer rank (engine/peer.py)


def _bound_pending():
    """Keep the pending list short when nobody calls check_finite: completed launches are consumed first (no
    stall), and past MAX_PENDING the oldest are waited for."""
    if len(_pending_flags) > MAX_PENDING:
        check_finite(block=False)
    if len(_pending_flags) > MAX_PENDING:
        old = _pending_flags[:len(_pending_flags) - MAX_PENDING]
        del _pending_flags[:len(old)]
        failed = None
        for names, flags, ev in old:
            for n, b in zip(names, flags.tolist()):
                if b != 0 and failed is None:
                    failed = n
        assert failed is None, "{} got nan or inf".format(failed)


class _Terms(torch.autograd.Function):
    """args: flat list of tensors referenced by `spec` (so autograd sees them)."""

    @staticmethod
    def forward(ctx, spec, nslots, *tensors):
        dev = tensors[0].device
        st = torch.cuda.current_stream(dev).cuda_stream
        segs = (N.ElboSeg * len(spec))()
        z_outs, names = [], []
        for i, s in enumerate(spec):
            a = tensors[s["a"]] if s.get("a") is not None else None
            b = tensors[s["b"]] if s.get("b") is not None else None
            out = None
            if s["kind"] == 1 and s.get("want_z", False):
                B, Z2, H, W = b.shape
                out = torch.empty((B, Z2 // 2, H, W), dtype=torch.float32, device=dev)
                z_outs.append(out)
            if s["kind"] == 1:
                B, Z2, H, W = b.shape
                Z, HW, n = Z2 // 2, H * W, B * (Z2 // 2) * H * W
            else:
                Z, HW, n = 0, 0, a.numel()
            segs[i] = N.ElboSeg(kind=s["kind"], slot=s["slot"], a=a.data_ptr() if a is not None else None,
                                b=b.data_ptr() if b is not None else None,
                                out=out.data_ptr() if out is not None else None, target=s.get("target", 0.0),
                                scale=s["scale"], Z=Z, HW=HW, n=n, prior=1 if s.get("prior") else 0)
            names.append(s.get("name", str(i)))
        table = _table(segs, dev)
        acc = torch.zeros(N.lib().vae2_elbo_acc_floats(), dtype=torch.float32, device=dev)
        flags = torch.zeros(len(spec), dtype=torch.int32, device=dev)
        N.call.vae2_elbo_terms(table.data_ptr(), len(spec), acc.data_ptr(), nslots, flags.data_ptr(), st)
        LAUNCHES["n"] += 2
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        _pending_flags.append((names, flags, ev))
        _bound_pending()
        ctx.spec, ctx.nslots = spec, nslots
        ctx.save_for_backward(*tensors)
        ctx.n_z = len(z_outs)
        ctx.keep = (table,)
        return (acc[:nslots].clone(),) + tuple(z_outs)

    @staticmethod
    def backward(ctx, gvals, *gz):
        tensors = ctx.saved_tensors
        spec = ctx.spec
        dev = tensors[0].device
        st = torch.cuda.current_stream(dev).cuda_stream
        gvals = torch.zeros(ctx.nslots, dtype=torch.float32, device=dev) if gvals is None else gvals.contiguous().float()
        grads = [None] * len(tensors)
        segs = (N.ElboBwdSeg * len(spec))()
        zi, n = 0, 0
        for s in spec:
            gidx = s["b"] if s["kind"] == 1 else s["a"]
            this_gz = None
            if s["kind"] == 1 and s.get("want_z", False):
                this_gz = gz[zi]
                zi += 1
            if not ctx.needs_input_grad[2 + gidx]:
                continue
            tgt = tensors[gidx]
            if grads[gidx] is None:
                grads[gidx] = torch.zeros_like(tgt)
            a = tensors[s["a"]] if s.get("a") is not None else None
            b = tensors[s["b"]] if s.get("b") is not None else None
            if s["kind"] == 1:
                B, Z2, H, W = b.shape
                Z, HW, cnt = Z2 // 2, H * W, B * (Z2 // 2) * H * W
                if this_gz is not None:
                    this_gz = this_gz.contiguous().float()
                    ctx.keep += (this_gz,)
            else:
                Z, HW, cnt = 0, 0, a.numel()
            segs[n] = N.ElboBwdSeg(kind=s["kind"], a=a.data_ptr() if a is not None else None,
                                   b=b.data_ptr() if b is not None else None,
                                   gz=this_gz.data_ptr() if this_gz is not None else None,
                                   grad=grads[gidx].data_ptr(), gout=gvals.data_ptr() + 4 * s["slot"],
                                   target=s.get("target", 0.0), scale=s["scale"], Z=Z, HW=HW, n=cnt, accumulate=1,
                                   prior=1 if s.get("prior") else 0)
            n += 1
        if n:
            table = _table(segs, dev)
            N.call.vae2_elbo_terms_bwd(table.data_ptr(), n, st)
            LAUNCHES["n"] += 1
            ctx.keep += (table, gvals)
        return (None, None) + tuple(grads)


def elbo_terms(spec, nslots, tensors):
    """spec: list of dicts {kind, slot, a, b, scale, target?, prior?, want_z?, name?} whose a/b are
    indices into `tensors` (contiguous fp32 CUDA).  Returns (vals[nslots], *z_outs)."""
    tensors = [t.contiguous().float() for t in tensors]
    if tensors[0].device.type != "cuda":
        raise RuntimeError("vae2_b200: ELBO kernels run on CUDA tensors only; there is no CPU fallback")
    return _Terms.apply(spec, nslots, *tensors)
