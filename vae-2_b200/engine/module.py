"""nn.Module <-> Plan glue: plan cache, autograd node, precision / graph switches.

An ``EngineModule`` keeps the reference's parameter tree (so ``state_dict`` keys, SyncBN
conversion, DDP hooks and optimizers all see ordinary nn.Parameters) but executes
``forward`` as ONE autograd node that replays a recorded Plan (engine/graph.py).
"""
import os
import weakref

import torch
import torch.distributed as dist
import torch.nn as nn

from .graph import ActArena, Plan, Recorder, activation_phase  # noqa: F401

_STATE = {
    "precision": os.environ.get("VAE2_PRECISION", "fp32"),
    "cuda_graphs": os.environ.get("VAE2_CUDA_GRAPHS", "0") == "1",
    "launches": 0,
}


def set_precision(name):
    """'fp32' (exact-fp32 storage and FMA) or 'bf16' (bf16 storage, tensor-core convs)."""
    assert name in ("fp32", "bf16")
    _STATE["precision"] = name


def get_precision():
    return _STATE["precision"]


def use_cuda_graphs(flag):
    _STATE["cuda_graphs"] = bool(flag)


def launch_count():
    """Launches of this library's kernels so far, replayed graph nodes included (bench.py's gpu_launches)."""
    from .elbo import LAUNCHES
    return _STATE["launches"] + LAUNCHES["n"]


class _Lease:
    """Marks a plan busy between forward and the end of backward (or graph destruction)."""

    def __init__(self, plan):
        self.plan = plan
        plan.busy = True
        if plan.arena_phase is not None:
            ActArena.live[plan.arena_phase] = ActArena.live.get(plan.arena_phase, 0) + 1

    def release(self):
        if self.plan is not None:
            self.plan.busy = False
            if self.plan.arena_phase is not None:
                ph = self.plan.arena_phase
                ActArena.live[ph] = max(0, ActArena.live.get(ph, 0) - 1)
            self.plan = None

    def __del__(self):
        self.release()


class _PlanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, n_in, *args):
        inputs = [a.detach().contiguous().float() for a in args[:n_in]]
        plan.cur_inputs = inputs
        dev = plan.device
        outs = [torch.empty((plan.B,) + tuple(s), dtype=torch.float32, device=dev) for s in plan.out_shapes]
        plan.cur_outputs = outs
        _STATE["launches"] += plan.run_forward(_STATE["cuda_graphs"])
        ctx.plan, ctx.n_in, ctx.n_args = plan, n_in, len(args)
        ctx.in_shapes = [tuple(a.shape) for a in args[:n_in]]
        ctx.in_needs = [bool(a.requires_grad) for a in args[:n_in]]
        if plan.training:   # plan.training already implies grad mode at the call site (grad mode is off in here)
            ctx.lease = _Lease(plan)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        plan = ctx.plan
        if not plan.training:
            raise RuntimeError("vae2_b200: backward through an eval-mode plan is not supported")
        dev = plan.device
        plan.cur_output_grads = [None if g is None else g.contiguous().float() for g in gouts]
        plan.cur_input_grads = [torch.zeros(s, dtype=torch.float32, device=dev) if need else None
                                for s, need in zip(ctx.in_shapes, ctx.in_needs)]
        _STATE["launches"] += plan.run_backward(_STATE["cuda_graphs"])
        flat = plan.flat_grad.clone()
        pgrads = []
        for p, off in zip(plan.params, plan._grad_off):
            pgrads.append(flat[off:off + p.numel()].view(p.shape) if p.requires_grad else None)
        in_grads = plan.cur_input_grads
        plan.cur_output_grads = plan.cur_input_grads = None
        lease = getattr(ctx, "lease", None)
        if lease is not None:
            lease.release()
        return (None, None) + tuple(in_grads) + tuple(pgrads)


class EngineModule(nn.Module):
    """Base for the module mirrors.  Subclasses implement

        _record(self, rec: Recorder, in_shapes, in_needs_grad) -> list of output (C,H,W)

    declaring inputs with ``rec.input``, emitting ops, and declaring outputs with ``rec.output``.
    """

    def _plans(self):
        d = self.__dict__.get("_plan_cache")
        if d is None:
            d = {}
            self.__dict__["_plan_cache"] = d   # bypass nn.Module attribute registration
        return d

    def reset_plans(self):
        for pool in self._plans().values():
            for plan in pool:
                if not plan.busy:
                    plan.release()
        self.__dict__["_plan_cache"] = {}

    def _sync_world(self):
        if dist.is_available() and dist.is_initialized():
            if any(isinstance(m, nn.SyncBatchNorm) for m in self.modules()):
                return dist.get_world_size()
        return 1

    def _run(self, inputs, tag="", batch=None, stat_groups=1):
        """``batch`` / ``stat_groups``: a stacked plan holds ``stat_groups`` calls of the reference along the batch
        axis (``batch`` samples in total), each normalised with its own BatchNorm statistics (graph.BnOp)."""
        dev = inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("vae2_b200: this path runs on CUDA devices only (got %s); there is no CPU fallback"
                               % dev)
        training = self.training
        grad_on = torch.is_grad_enabled()
        needs = tuple(bool(t.requires_grad) and grad_on for t in inputs)
        shapes = tuple(tuple(t.shape) for t in inputs)
        world = self._sync_world() if training else 1
        frozen = tuple(not p.requires_grad for p in self.parameters()) if (training and grad_on) else ()
        frozen = hash(frozen) if any(frozen) else 0
        key = (tag, shapes, needs, training, grad_on, _STATE["precision"], dev.index, world, frozen, stat_groups)
        pool = self._plans().setdefault(key, [])
        plan = next((p for p in pool if not p.busy), None)
        if plan is not None and plan.param_ptrs != tuple(p.data_ptr() for p in plan.params):
            for old in pool:  # parameters were re-allocated (.to(), .cuda(), ...): recorded pointers are stale
                if not old.busy:
                    old.release()
            pool.clear()
            plan = None
        if plan is None:
            with torch.cuda.device(dev):
                plan = Plan(dev, batch or shapes[0][0], _STATE["precision"], training and grad_on, bn_batch_stats=training,
                            world_size=world, stat_groups=stat_groups)
                rec = Recorder(plan)
                plan.out_shapes = self._record(rec, shapes, needs, tag)
                plan.finalize()
                plan.param_ptrs = tuple(p.data_ptr() for p in plan.params)
            pool.append(plan)
        if plan.arena_phase is not None:
            for other, n in ActArena.live.items():
                if other != plan.arena_phase and n > 0 and ActArena.region(other) == ActArena.region(plan.arena_phase):
                    raise RuntimeError("vae2_b200: a network of phase %r still waits for its backward while phase %r runs; "
                                       "the phases share activation memory (set VAE2_ACT_ARENA=0 to interleave them)"
                                       % (other, plan.arena_phase))
        params = [p for p in plan.params]
        with torch.cuda.device(dev):
            outs = _PlanFn.apply(plan, len(inputs), *inputs, *params)
        return outs
