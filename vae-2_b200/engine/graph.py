"""Static execution plans for the VAE^2 networks.

A network (or a single block) is *recorded* once per (input shapes, mode) into a Plan: a
flat list of kernel launches over pre-allocated channels-last buffers.  Forward and backward
are then replays of that list on one CUDA stream -- eagerly, or as a captured CUDA graph --
with no per-op Python/autograd work on the hot path.  This replaces the reference's
op-by-op nn.Module execution (lib/models/enc_hrnet.py forward methods) while keeping its
parameters, state_dict and autograd-visible inputs/outputs.

Only plumbing uses torch here (device memory, streams, torch.distributed for SyncBN
all-gathers); every arithmetic op is a call into libvae2_b200.so through the C ABI.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import native as N
from . import peer

BN_EPS_DEFAULT = 1e-5


class Precision:
    def __init__(self, name, code, tdtype, seg_align, tot_align):
        self.name, self.code, self.tdtype = name, code, tdtype
        self.seg_align, self.tot_align = seg_align, tot_align
        self.esize = 4 if code == 0 else 2


PRECISIONS = {
    "fp32": Precision("fp32", 0, torch.float32, 4, 4),
    "bf16": Precision("bf16", 1, torch.bfloat16, 8, 16),
}


def pad_to(c, a):
    return (c + a - 1) // a * a


def conv_out(n, k, s):
    p = k // 2
    return (n + 2 * p - k) // s + 1


class GradArena:
    """One gradient arena per (device, dtype), shared by ALL plans: activation gradients only live inside
    a plan's run_backward (written from the upstream gradient in pre_bwd, consumed by post_bwd), and
    backward passes run one at a time, so every plan can alias the same memory.  This halves the
    resident footprint of a training step (the D nets alone are 12 plan instances per iteration)."""
    _arenas = {}

    @classmethod
    def get(cls, device, tdtype, numel):
        key = (str(device), tdtype)
        cur = cls._arenas.get(key)
        if cur is None or cur.numel() < numel:
            cur = torch.zeros(numel, dtype=tdtype, device=device)   # older plans keep their (smaller) arena alive
            cls._arenas[key] = cur
        return cur

    @classmethod
    def reset(cls):
        cls._arenas.clear()


class ActArena:
    """Activation memory shared between the PHASES of a training iteration.

    The generator step (FullModel_encdec: encz + encdec + 4 discriminator passes) and the discriminator step
    (FullModel_D: 8 discriminator passes) never overlap in time: each runs forward -> backward to completion.  Plans
    recorded while a phase is active bump-allocate their activation buffers from a chunk list that every phase walks
    from the start, so phase "D" lives in the memory phase "G" used a moment ago; the resident footprint is the
    maximum over phases instead of the sum (113 -> ~75 GB at B=2 fp32).  Within a phase all plans are disjoint.
    Safety: a forward of an arena-backed plan while a plan of ANOTHER phase still waits for its backward raises
    (EngineModule._run), and buffers that no kernel rewrites completely (concat roots: the lanes behind the last
    segment) are re-zeroed at the start of every forward.  Outside a phase (direct module calls, tests) plans get
    private memory as before.  VAE2_ACT_ARENA=0 disables the sharing."""
    CHUNK_BYTES = 4 << 30        # an allocation that does not fit the rest of a chunk starts the next one: <= ~7 % waste
    phase = None
    live = {}
    _pools = {}

    @staticmethod
    def region(phase):
        """Phases named 'region:name' live in their own pool: 'scratch:*' holds the transient discriminator passes whose
        backward runs right after their forward (utils.utils._EagerGan*), next to -- not on top of -- the 'main' pool
        in which the generator networks wait for their backward."""
        return phase.split(":", 1)[0] if (phase is not None and ":" in phase) else "main"

    @classmethod
    def enabled(cls):
        return os.environ.get("VAE2_ACT_ARENA", "1") != "0"

    @classmethod
    def alloc(cls, device, tdtype, numel, owner=None):
        """Bump-allocate from the current phase's cursor, after a first-fit look at the extents that dropped plans
        of this phase gave back (Plan.release): re-recording a plan (new batch size, parameters moved, reset_plans)
        reuses its predecessor's memory instead of growing the arena."""
        if cls.phase is None or not cls.enabled():
            return None
        esize = torch.empty(0, dtype=tdtype).element_size()
        key = (str(device), tdtype, cls.region(cls.phase))
        pool = cls._pools.setdefault(key, {"chunks": [], "cur": {}, "free": {}})
        n = pad_to(numel, 128)
        free = pool["free"].setdefault(cls.phase, [])
        for i, (ci, off, sz) in enumerate(free):
            if sz >= n:
                if sz == n:
                    free.pop(i)
                else:
                    free[i] = (ci, off + n, sz - n)
                return cls._take(pool, key, ci, off, n, numel, owner)
        cur = pool["cur"].setdefault(cls.phase, [0, 0])
        while True:
            ci, off = cur
            if ci >= len(pool["chunks"]):
                pool["chunks"].append(torch.empty(max(n, cls.CHUNK_BYTES // esize), dtype=tdtype, device=device))
            chunk = pool["chunks"][ci]
            if off + n <= chunk.numel():
                cur[1] = off + n
                return cls._take(pool, key, ci, off, n, numel, owner)
            if chunk.numel() - off >= 128:      # the tail of a chunk that did not fit stays usable for smaller buffers
                free.append((ci, off, chunk.numel() - off))
            cur[0], cur[1] = ci + 1, 0

    @classmethod
    def _take(cls, pool, key, ci, off, n, numel, owner):
        t = pool["chunks"][ci][off:off + numel]
        t.zero_()
        if owner is not None:
            owner.arena_extents.append((key, cls.phase, ci, off, n))
        return t

    @classmethod
    def give_back(cls, extents):
        for key, phase, ci, off, n in extents:
            pool = cls._pools.get(key)
            if pool is None or ci >= len(pool["chunks"]):
                continue                 # the arena was reset in the meantime
            free = pool["free"].setdefault(phase, [])
            free.append((ci, off, n))
        for pool in cls._pools.values():  # coalesce neighbours
            for phase, free in pool["free"].items():
                free.sort()
                out = []
                for e in free:
                    if out and out[-1][0] == e[0] and out[-1][1] + out[-1][2] == e[1]:
                        out[-1] = (e[0], out[-1][1], out[-1][2] + e[2])
                    else:
                        out.append(e)
                free[:] = out

    @classmethod
    def reset(cls):
        cls._pools.clear()
        cls.live.clear()

    @classmethod
    def summary(cls):
        """{region: GiB held} plus the gradient arenas -- where a step's resident memory sits (bench.py reports it)."""
        out = {}
        for (dev, tdt, region), pool in cls._pools.items():
            out[region] = out.get(region, 0.0) + sum(c.numel() * c.element_size() for c in pool["chunks"]) / 2**30
        out["grad_arena"] = sum(t.numel() * t.element_size() for t in GradArena._arenas.values()) / 2**30
        return {k: round(v, 2) for k, v in out.items()}


class activation_phase:
    """Context manager used by the wrapper modules: plans recorded inside share activation memory across phases."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.prev = ActArena.phase
        ActArena.phase = self.name

    def __exit__(self, *exc):
        ActArena.phase = self.prev
        return False


class _FakeBuf:
    """Stand-in for a gradient buffer during the sizing pass of Plan._emit_backward (only addresses are taken)."""

    def __init__(self, byte_off, numel):
        self._p, self._n = 4096 + byte_off, numel

    def data_ptr(self):
        return self._p

    def numel(self):
        return self._n

    def zero_(self):
        return self


class Act:
    """A channels-last activation [B][H][W][ld] (possibly a channel slice of a wider root)."""

    def __init__(self, plan, C_, H, W, root=None, c_off=0, Cp=None, name="", B=None):
        self.plan, self.C, self.H, self.W = plan, C_, H, W
        self.B = (root.B if root is not None else plan.B) if B is None else B   # sub-batch acts: K-sample inference
        self.name = name
        pr = plan.prec
        if root is None:
            self.Cp = pad_to(C_, pr.tot_align) if Cp is None else Cp
            self.numel = self.B * H * W * self.Cp
            if plan.lazy_acts:
                self._buf = None          # inference plan: assigned by Plan._assign_buffers from buffer liveness
            else:
                self._buf = ActArena.alloc(plan.device, pr.tdtype, self.numel, owner=plan)
                if self._buf is None:
                    self._buf = torch.zeros(self.numel, dtype=pr.tdtype, device=plan.device)
                else:
                    plan.arena_phase = ActArena.phase
            self.ld, self.c_off, self.parent = self.Cp, 0, None
            plan.all_acts.append(self)
        else:
            self.Cp = Cp
            self._buf, self.ld, self.c_off, self.parent = None, root.ld, root.c_off + c_off, root
        self._grad = None
        self.needs_grad = plan.training
        self.cin_map = None      # logical channel -> physical lane, for concat roots
        self.grad_written = False

    @property
    def npix(self):
        return self.B * self.H * self.W

    @property
    def buf(self):
        a = self
        while a.parent is not None:
            a = a.parent
        return a._buf

    @property
    def ptr(self):
        return self.buf.data_ptr() + self.c_off * self.plan.prec.esize

    @property
    def root(self):
        a = self
        while a.parent is not None:
            a = a.parent
        return a

    def grad(self):
        """Gradient buffer with the same geometry (a slice of the root's gradient for slices)."""
        if self._grad is None:
            if self.parent is None:
                g = Act.__new__(Act)
                g.__dict__.update(self.__dict__)
                g._buf, g._gext = self.plan.grad_alloc(self.numel)
                g._grad, g.parent = None, None
                self._grad = g
            else:
                pg = self.parent.grad()
                g = Act.__new__(Act)
                g.__dict__.update(self.__dict__)
                g._buf, g.parent, g._grad = None, pg, None
                self._grad = g
        return self._grad

    # gradient "first write vs accumulate" bookkeeping lives on the root buffer
    def take_acc_flag(self):
        r = self.root
        acc = r.grad_written
        r.grad_written = True
        return 1 if acc else 0

    def to_nchw(self):
        """Debug/test helper: logical NCHW fp32 copy of this activation."""
        v = self.buf[:self.root.numel].view(self.B, self.H, self.W, self.ld)[..., self.c_off:self.c_off + self.C]
        return v.permute(0, 3, 1, 2).float().contiguous()


class Plan:
    """Recorded launch list + buffers for one network call signature."""

    def __init__(self, device, B, prec, training, bn_batch_stats=None, world_size=1, process_group=None, stat_groups=1):
        self.device, self.B, self.prec, self.training = device, B, PRECISIONS[prec], training
        self.stat_groups = stat_groups        # BN statistics groups along the batch (stacked calls of the reference)
        # Inference plans (no backward program) keep an activation only while a later op still reads it: buffers are
        # assigned from a free list after recording (_assign_buffers).  K-sample inference at K=16 would otherwise hold
        # every intermediate of 16 samples.  VAE2_LAZY_ACTS=0 restores one private buffer per activation.
        self.lazy_acts = (not training) and os.environ.get("VAE2_LAZY_ACTS", "1") != "0"
        # training: record a backward program.  bn_batch_stats: BN normalises with batch statistics
        # (module.training), which also holds for a train-mode forward under torch.no_grad().
        self.bn_batch_stats = training if bn_batch_stats is None else bn_batch_stats
        self.world_size, self.group = world_size, process_group
        self.fwd, self.bwd = [], []          # lists of zero-arg callables taking (stream_ptr)
        self.pre_fwd, self.post_fwd = [], []  # eager: input conversion / output conversion
        self.pre_bwd, self.post_bwd = [], []
        self.ops = []
        self.params = []                      # nn.Parameters in registration order
        self._param_index = {}
        self._grad_off = []
        self.flat_grad = None
        self._wp_sizes, self._wpT_sizes = [], []
        self.wp_flat = self.wpT_flat = self.dwp_flat = None
        self.pack_descs = []                  # filled at finalize
        self.inputs, self.outputs = [], []
        self.keep = []                        # keep-alive for ctypes arrays / tensors
        self.busy = False
        self.n_collectives_fwd = self.n_collectives_bwd = 0
        self._garena, self._goff, self._gsim, self._gfree, self._gpeak = None, 0, False, [], 0
        self.all_acts = []
        self.arena_phase = None               # set when an activation buffer comes from the phase-shared ActArena
        self.arena_extents = []               # (pool key, phase, chunk, offset, size) of those buffers
        self.released = False
        self.taps = {}                        # name -> Act at module boundaries (debug / parity tests: Act.to_nchw)
        self.concat_roots = []
        self.use_tc = os.environ.get("VAE2_DISABLE_TC", "0") != "1"
        # fp32 storage, forward / data-gradient GEMMs of the >= 40-lane layers on tcgen05 through the exact 3-way bf16 split
        # with split accumulators (csrc/conv_f32x3.cu): measured MORE accurate than the CUDA-core FMA loops inside the
        # full nets (profiles/r2_parity_errors.jsonl: x2p 1.6e-6 vs 2.8e-6 on tiny, 7.2e-5 vs 1.06e-4 on W48).
        # "0" = CUDA cores only, "all" = every supported conv (the narrow 20/36-lane layers are faster on conv_direct).
        self.fp32_tc = os.environ.get("VAE2_FP32_TC", "1")
        self.fp32_tc_min_lanes = int(os.environ.get("VAE2_FP32_TC_MIN_LANES", "40"))
        self.use_tc_wgrad = os.environ.get("VAE2_DISABLE_TC_WGRAD", "0") != "1"
        self.graph_fwd = self.graph_bwd = None
        self.n_launch_fwd = self.n_launch_bwd = 0

    # ---- registration ------------------------------------------------------------------
    def param(self, p):
        key = id(p)
        if key not in self._param_index:
            self._param_index[key] = len(self.params)
            self.params.append(p)
        return self._param_index[key]

    def new_act(self, C_, H, W, name="", B=None):
        return Act(self, C_, H, W, name=name, B=B)

    def grad_alloc(self, numel):
        """An activation-gradient buffer from the shared arena (128-element aligned): first fit among the extents that
        finished gradients gave back (_release_grads), else the bump cursor.  Returns (buffer, extent).  During the
        sizing pass (_gsim) the buffer is a stand-in that only knows its offset."""
        if self._garena is None and not self._gsim:      # VAE2_PRIVATE_GRADS=1 / a stray early request: private memory
            return torch.zeros(numel, dtype=self.prec.tdtype, device=self.device), None
        n = pad_to(numel, 128)
        off = None
        for i, (o, sz) in enumerate(self._gfree):
            if sz >= n:
                off = o
                if sz == n:
                    self._gfree.pop(i)
                else:
                    self._gfree[i] = (o + n, sz - n)
                break
        if off is None:
            off = self._goff
            self._goff += n
            self._gpeak = max(self._gpeak, self._goff)
        if self._gsim:
            return _FakeBuf(off * self.prec.esize, numel), (off, n)
        assert off + n <= self._garena.numel(), "gradient arena under-sized"
        return self._garena[off:off + numel], (off, n)

    def _release_grads(self, op):
        """After op's backward has been emitted: the gradient of every buffer whose producers are all done is dead."""
        for a in op.writes():
            r = a.root
            r._nw -= 1
            g = r._grad
            if r._nw == 0 and g is not None and getattr(g, "_gext", None) is not None and not r._pin:
                self._gfree.append(g._gext)
                g._gext = None
                self._gfree.sort()
                out = []
                for e in self._gfree:           # coalesce neighbours
                    if out and out[-1][0] + out[-1][1] == e[0]:
                        out[-1] = (out[-1][0], out[-1][1] + e[1])
                    else:
                        out.append(e)
                self._gfree = out

    def _emit_backward(self, convs):
        """(Re)build the backward program.  Gradient buffers are recycled: a gradient lives from the backward of the
        buffer's last consumer to the backward of its first producer, so the arena holds the peak over the program,
        not one buffer per activation (63 GB -> a few GB for a stacked discriminator pass at B=4)."""
        self.bwd, self.pre_bwd, self.post_bwd = [], [], []
        self.n_collectives_bwd = 0
        self._gfree, self._goff, self._gpeak = [], 0, 0
        acts = {id(a): a for a in self.all_acts}
        for o in self.ops:
            for a in list(o.reads()) + list(o.writes()):
                acts[id(a)] = a
        for a in acts.values():
            a._grad, a.grad_written, a._nw, a._pin = None, False, 0, False
        for o in self.ops:
            for a in o.writes():
                a.root._nw += 1
            if isinstance(o, (InputOp, GroupInputOp, CodeOp)) and o.needs_grad:
                for a in o.writes():
                    a.root._pin = True      # read by the eager epilogue AFTER the whole program: never recycled
            if isinstance(o, OutputOp):
                o.act.grad()                # written by the eager prologue BEFORE the program: allocated first
        dptr = self.dwp_flat
        any_wgrad = any(o.conv.weight.requires_grad for o in convs)
        if any_wgrad:
            self.bwd.append(lambda st: dptr.zero_())
        for o in reversed(self.ops):
            o.emit_bwd(self)
            self._release_grads(o)
        if any_wgrad:
            self.bwd.append(lambda st: N.call.vae2_unpack_wgrad(self.up_dev.data_ptr(), self.n_convs, 0, st))

    def concat(self, Cs, H, W, name="cat", B=None):
        """A buffer made of padded channel segments; returns (root, [slice acts])."""
        pr = self.prec
        seg_p = [pad_to(c, pr.seg_align) for c in Cs]
        total = pad_to(sum(seg_p), pr.tot_align)
        root = Act(self, sum(Cs), H, W, Cp=total, name=name, B=B)
        self.concat_roots.append(root)
        cmap, off, slices = [], 0, []
        for c, cp in zip(Cs, seg_p):
            slices.append(Act(self, c, H, W, root=root, c_off=off, Cp=cp, name="%s[%d]" % (name, off)))
            cmap += list(range(off, off + c))
            off += cp
        root.cin_map = cmap
        return root, slices

    def add(self, op):
        self.ops.append(op)
        return op

    # ---- build -------------------------------------------------------------------------
    def finalize(self):
        dev = self.device
        # flat parameter-gradient arena
        off = 0
        for p in self.params:
            self._grad_off.append(off)
            off += pad_to(p.numel(), 4)
        self.flat_grad = torch.zeros(max(off, 4), dtype=torch.float32, device=dev)
        # packed weight arenas
        convs = [o for o in self.ops if isinstance(o, ConvOp)]
        tot = 0
        for o in convs:
            o.w_off = tot
            tot += pad_to(o.taps * o.x.root_cp() * o.y.Cp, 4)
        # tensor-core eligibility: decided once per conv at plan build
        if self.prec.code == 1 and dev.type == "cuda" and self.use_tc:
            for o in convs:
                g = o._geom()
                o.engine = 1 if (o.x.H * o.x.W > 1 and N.lib().vae2_conv2d_tc_supported(C.byref(g))) else 0
        # fp32 storage with the fwd/dgrad GEMMs on tensor cores through the exact 3-way bf16 split (opt-in:
        # its truncating fp32 accumulation is ~1e-6..1e-5 per conv)
        # which layers: >= 40 lanes on either side, and the 3x3 stride-1 layers from 36 lanes up (they take the halo-tile
        # kernel: measured 36->36 0.40 vs 0.60 ms on conv_direct at B=24, but 18->18 0.78 vs 0.65 ms -- N=32 MMAs cost
        # what N=64 ones do)
        hmin = int(os.environ.get("VAE2_FP32_TC_HALO_MIN_LANES", "36"))

        def on_tc(o):
            if self.fp32_tc == "all" or max(o.x.root_cp(), o.y.Cp) >= self.fp32_tc_min_lanes:
                return True
            return o.conv.kernel_size[0] == 3 and o.conv.stride[0] == 1 and min(o.x.root_cp(), o.y.Cp) >= hmin
        x3 = [o for o in convs if self.prec.code == 0 and self.fp32_tc not in ("0", "") and dev.type == "cuda"
              and on_tc(o) and o.x.H * o.x.W > 1 and N.lib().vae2_conv2d_tf32_supported(C.byref(o._geom()))]
        tot3f = tot3b = 0
        for o in x3:
            o.engine = 2
            o.x3dims = N.tf32_dims(o._geom())
            Nf, Kf, NfT, KfT = o.x3dims
            o.x3_foff, o.x3_boff = tot3f, tot3b
            tot3f += pad_to(3 * o.taps * Nf * Kf, 8)
            tot3b += pad_to(3 * o.taps * NfT * KfT, 8)
        self.w3f_flat = torch.zeros(max(tot3f, 8), dtype=torch.bfloat16, device=dev) if x3 else None
        self.w3b_flat = torch.zeros(max(tot3b, 8), dtype=torch.bfloat16, device=dev) if (x3 and self.training) else None
        if x3:
            t3 = (N.Tf32PackDesc * len(x3))()
            for i, o in enumerate(x3):
                w = o.conv.weight
                cm = None
                if o.x.cin_map is not None:
                    t = torch.tensor(o.x.cin_map, dtype=torch.int32, device=dev)
                    self.keep.append(t)
                    cm = t.data_ptr()
                Nf, Kf, NfT, KfT = o.x3dims
                t3[i] = N.Tf32PackDesc(w=w.data_ptr(), cin_map=cm, fwd=self.w3f_flat.data_ptr() + 2 * o.x3_foff,
                                       bwd=(self.w3b_flat.data_ptr() + 2 * o.x3_boff) if self.training else None,
                                       Cout=w.shape[0], Cin=w.shape[1], k=w.shape[2], Nf=Nf, Kf=Kf, NfT=NfT, KfT=KfT)
            self.t3_dev = torch.frombuffer(bytearray(bytes(t3)), dtype=torch.uint8).to(dev)
            n3 = len(x3)
            self.fwd.append(lambda st: N.call.vae2_pack_weights_tf32(self.t3_dev.data_ptr(), n3, st))
        self.n_tc_convs = sum(o.engine for o in convs)
        # shared split-K workspace of the tensor-core weight gradient (convs run one after another)
        ws_floats = 0
        if self.training:
            for o in convs:
                o.wgrad_tc = False
                if o.engine == 1 and self.use_tc_wgrad:
                    g = o._geom()
                    need = N.lib().vae2_conv2d_wgrad_tc_workspace(C.byref(g))
                    if need > 0:
                        o.wgrad_tc = True
                        ws_floats = max(ws_floats, need)
        self.wgrad_ws = torch.zeros(ws_floats, dtype=torch.float32, device=dev) if ws_floats else None
        # fp32 path: weight gradient of the tensor-core (engine 2) layers through bf16 hi/lo planes on the tcgen05 wgrad
        # kernel (csrc/conv_tc.cu: conv_wgrad_f32x2); one shared workspace, the convs run one after another
        x2_bytes = 0
        if self.training and os.environ.get("VAE2_FP32_TC_WGRAD", "1") != "0":
            wmin = int(os.environ.get("VAE2_FP32_TC_WGRAD_MIN_LANES", "36"))
            # the 3x3 stride-1 layers of <= 32 input lanes (18 -> 18) as well: their planes go through the halo-tile
            # kernel in ONE pass (csrc/conv_tc.cu, dual mode) whatever engine runs their forward
            narrow = os.environ.get("VAE2_FP32_TC_WGRAD_NARROW", "1") != "0" and self.fp32_tc not in ("0", "")
            for o in convs:
                if self.prec.code != 0 or dev.type != "cuda" or o.x.H * o.x.W <= 1:
                    continue
                wide = o in x3 and max(o.x.root_cp(), o.y.Cp) >= wmin
                halo1 = (narrow and o.conv.kernel_size[0] == 3 and o.conv.stride[0] == 1 and 16 < o.x.root_cp() <= 32
                         and o.y.Cp <= 80)
                if not (wide or halo1):
                    continue
                need = N.lib().vae2_conv2d_wgrad_f32x2_workspace(C.byref(o._geom()))
                o.wgrad_x2 = need > 0 and o.conv.weight.requires_grad
                if o.wgrad_x2:
                    x2_bytes = max(x2_bytes, need)
        self.wgrad_x2_ws = torch.empty(x2_bytes, dtype=torch.uint8, device=dev) if x2_bytes else None
        self.wp_flat = torch.zeros(max(tot, 4), dtype=torch.float32, device=dev)
        self.wpT_flat = torch.zeros(max(tot, 4), dtype=torch.float32, device=dev) if self.training else None
        self.dwp_flat = torch.zeros(max(tot, 4), dtype=torch.float32, device=dev) if self.training else None
        any_tc = self.n_tc_convs > 0
        self.wq_flat = torch.zeros(max(tot, 8), dtype=torch.bfloat16, device=dev) if any_tc else None
        self.wqT_flat = torch.zeros(max(tot, 8), dtype=torch.bfloat16, device=dev) if (any_tc and self.training) else None
        # pack / unpack descriptor tables (device resident)
        pk = (N.PackDesc * max(len(convs), 1))()
        up = (N.PackDesc * max(len(convs), 1))()
        for i, o in enumerate(convs):
            w = o.conv.weight
            cmap_ptr = None
            if o.x.cin_map is not None:
                t = torch.tensor(o.x.cin_map, dtype=torch.int32, device=dev)
                self.keep.append(t)
                cmap_ptr = t.data_ptr()
            cin_p, cout_p = o.x.root_cp(), o.y.Cp
            common = dict(cin_map=cmap_ptr, Cout=w.shape[0], Cin=w.shape[1], k=w.shape[2], Cin_p=cin_p, Cout_p=cout_p)
            if o.engine == 1:
                pk[i] = N.PackDesc(w=w.data_ptr(), wq=self.wq_flat.data_ptr() + 2 * o.w_off,
                                   wqT=(self.wqT_flat.data_ptr() + 2 * o.w_off) if self.training else None, **common)
            else:
                pk[i] = N.PackDesc(w=w.data_ptr(), wp=self.wp_flat.data_ptr() + 4 * o.w_off,
                                   wpT=(self.wpT_flat.data_ptr() + 4 * o.w_off) if self.training else None, **common)
            if self.training:
                gi = self.param(w)
                up[i] = N.PackDesc(w=self.flat_grad.data_ptr() + 4 * self._grad_off[gi],
                                   wp=self.dwp_flat.data_ptr() + 4 * o.w_off, **common)
        self._pk_host, self._up_host = pk, up
        self.pk_dev = torch.frombuffer(bytearray(bytes(pk)), dtype=torch.uint8).to(dev)
        self.up_dev = torch.frombuffer(bytearray(bytes(up)), dtype=torch.uint8).to(dev)
        self.n_convs = len(convs)
        # forward program
        self.fwd.append(lambda st: N.call.vae2_pack_weights(self.pk_dev.data_ptr(), self.n_convs, st))
        zero_at = self._assign_buffers() if self.lazy_acts else {}
        for i, o in enumerate(self.ops):
            for r in zero_at.get(i, ()):        # a recycled concat root: its never-written pad lanes must read as zero
                self.fwd.append(lambda st, b=r._buf, n=r.numel: b[:n].zero_())
            o.emit_fwd(self)
        if self.arena_phase is not None:
            # shared memory: the lanes of a concat root that no producer writes must not keep another phase's data.
            # First thing of the eager prologue, i.e. before the input/code ops write their slices.
            for r in self.concat_roots:
                self.pre_fwd.insert(0, lambda st, b=r.buf: b.zero_())
        # backward program (reverse order; accumulate flags resolved statically)
        if self.training:
            self._garena, self._gsim = None, False
            if os.environ.get("VAE2_PRIVATE_GRADS", "0") != "1":
                self._gsim = True                    # sizing pass: same emission, stand-in buffers, measures the peak
                self._emit_backward(convs)
                self._gsim = False
                self.grad_arena_elems = self._gpeak
                self._garena = GradArena.get(dev, self.prec.tdtype, max(self._gpeak, 128))
            self._emit_backward(convs)
        self.n_launch_fwd, self.n_launch_bwd = len(self.fwd), len(self.bwd)
        return self

    def _assign_buffers(self):
        """Liveness-based buffer assignment for inference plans.  Ops run in list order; inputs are written before op 0
        (eager prologue), outputs are read after the last op (eager epilogue).  A root buffer is taken from the free
        list (smallest that fits) when its first writer runs and returned after its last reader.  Returns
        {op index: [concat roots to zero before that op]}.  VAE2_KEEP_TAPS=1 keeps the tap activations private (debugging)."""
        first, last = {}, {}
        n = len(self.ops)

        def touch(act, i, write):
            r = act.root
            if write:
                first[id(r)] = min(first.get(id(r), i), i)
            last[id(r)] = max(last.get(id(r), i), i)

        for i, o in enumerate(self.ops):
            for a in o.writes():
                touch(a, -1 if isinstance(o, (InputOp, GroupInputOp, CodeOp)) else i, True)
            for a in o.reads():
                touch(a, n if isinstance(o, OutputOp) else i, False)
        keep = {id(a.root) for a in self.taps.values()} if os.environ.get("VAE2_KEEP_TAPS", "0") == "1" else set()
        roots = {id(a): a for a in self.all_acts}
        free, zero_at = [], {}
        by_first = {}
        for rid, a in roots.items():
            by_first.setdefault(first.get(rid, -1), []).append(a)
        by_last = {}
        for rid, a in roots.items():
            by_last.setdefault(last.get(rid, first.get(rid, -1)), []).append(a)
        tdt, dev = self.prec.tdtype, self.device
        concat = {id(r) for r in self.concat_roots}
        self.lazy_bytes = 0
        for i in range(-1, n + 1):
            for a in sorted(by_first.get(i, []), key=lambda t: -t.numel):
                cand = [b for b in free if b.numel() >= a.numel]
                if cand and id(a) not in keep and i >= 0:
                    b = min(cand, key=lambda t: t.numel())
                    free[:] = [t for t in free if t is not b]
                    if id(a) in concat:
                        zero_at.setdefault(i, []).append(a)
                else:
                    b = torch.zeros(a.numel, dtype=tdt, device=dev)
                    self.lazy_bytes += a.numel * self.prec.esize
                a._buf = b
            for a in by_last.get(i, []):
                if id(a) not in keep and first.get(id(a), -1) >= 0 and i < n:
                    free.append(a._buf)
        return zero_at

    def release(self):
        """Drop this plan: its arena extents go back to the phase's free list (the plan must not run again)."""
        self.released = True
        ActArena.give_back(self.arena_extents)
        self.arena_extents = []
        self.graph_fwd = self.graph_bwd = None

    def grad_view(self, p):
        i = self._param_index[id(p)]
        return self.flat_grad[self._grad_off[i]:self._grad_off[i] + p.numel()].view(p.shape)

    def grad_ptr(self, p):
        return self.flat_grad.data_ptr() + 4 * self._grad_off[self.param(p)]

    # ---- run ---------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def run_forward(self, use_graph=False):
        """Returns the number of native (C-ABI) launches this replay performed."""
        st = self._stream()
        c0 = N.COUNTERS["native_calls"]
        for f in self.pre_fwd:
            f(st)
        eager = N.COUNTERS["native_calls"] - c0
        if use_graph:
            if self.graph_fwd is None:
                # The warm-up pass inside _capture IS this call's execution (capturing launches nothing), so the
                # graph is not replayed on top of it: forward side effects (BN running statistics,
                # num_batches_tracked) must happen exactly once per call, as in eager mode and in the reference.
                c1 = N.COUNTERS["native_calls"]
                self.graph_fwd = self._capture(self.fwd)
                self.native_fwd = (N.COUNTERS["native_calls"] - c1) // 2   # warm-up pass + capture pass
            else:
                self.graph_fwd.replay()
            body = self.native_fwd
        else:
            c1 = N.COUNTERS["native_calls"]
            for f in self.fwd:
                f(st)
            body = N.COUNTERS["native_calls"] - c1
        c2 = N.COUNTERS["native_calls"]
        for f in self.post_fwd:
            f(st)
        return eager + body + N.COUNTERS["native_calls"] - c2

    def run_backward(self, use_graph=False):
        st = self._stream()
        c0 = N.COUNTERS["native_calls"]
        for f in self.pre_bwd:
            f(st)
        eager = N.COUNTERS["native_calls"] - c0
        if use_graph:
            if self.graph_bwd is None:
                c1 = N.COUNTERS["native_calls"]
                self.graph_bwd = self._capture(self.bwd)      # its warm-up pass is this call's execution
                self.native_bwd = (N.COUNTERS["native_calls"] - c1) // 2
            else:
                self.graph_bwd.replay()
            body = self.native_bwd
        else:
            c1 = N.COUNTERS["native_calls"]
            for f in self.bwd:
                f(st)
            body = N.COUNTERS["native_calls"] - c1
        c2 = N.COUNTERS["native_calls"]
        for f in self.post_bwd:
            f(st)
        return eager + body + N.COUNTERS["native_calls"] - c2

    def _capture(self, prog):
        # warm-up run on a side stream, then capture the same launch list
        torch.cuda.synchronize(self.device)
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for f in prog:
                f(s.cuda_stream)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in prog:
                f(s.cuda_stream)
        return g


def _root_cp(self):
    return self.Cp


Act.root_cp = _root_cp


# =============================================================================================
# ops
# =============================================================================================
class InputOp:
    """External NCHW fp32 tensor (channel window) -> channels-last activation(s).  Eager."""

    def reads(self):
        return []

    def writes(self):
        return list(self.dsts)

    def __init__(self, plan, slot, C_, H, W, src_ctot, src_coff, dsts, needs_grad):
        self.slot, self.C, self.H, self.W = slot, C_, H, W
        self.src_ctot, self.src_coff, self.dsts, self.needs_grad = src_ctot, src_coff, dsts, needs_grad
        for d in dsts:
            d.needs_grad = needs_grad and plan.training

    def emit_fwd(self, plan):
        pr = plan.prec

        def run(st, self=self):
            src = plan.cur_inputs[self.slot]
            for d in self.dsts:
                N.call.vae2_nchw_to_act(src.data_ptr(), d.ptr, pr.code, d.B, self.C, d.Cp, self.H, self.W, d.ld,
                                        self.src_ctot, self.src_coff, st)
        plan.pre_fwd.append(run)

    def emit_bwd(self, plan):
        if not (self.needs_grad and plan.training):
            return
        pr = plan.prec

        def run(st, self=self):
            dst = plan.cur_input_grads[self.slot]
            for i, d in enumerate(self.dsts):
                N.call.vae2_act_to_nchw(d.grad().ptr, dst.data_ptr(), pr.code, d.B, self.C, self.H, self.W, d.ld,
                                        self.src_ctot, self.src_coff, 1, st)   # dst starts zero-filled
        plan.post_bwd.append(run)


class GroupInputOp:
    """Statistics group `gi` of a stacked plan: samples [gi*Bg, (gi+1)*Bg) of `dst` come from a channel window of
    external NCHW tensor #slot (the reference calls the network once per group, e.g. per frame of the predicted clip,
    lib/utils/utils.py:116-119, 262-267).  Eager."""

    def reads(self):
        return []

    def writes(self):
        return [self.dst]

    def __init__(self, plan, slot, C_, H, W, src_ctot, src_coff, dst, gi, Bg, needs_grad):
        self.slot, self.C, self.H, self.W = slot, C_, H, W
        self.src_ctot, self.src_coff, self.dst, self.gi, self.Bg = src_ctot, src_coff, dst, gi, Bg
        self.needs_grad = needs_grad and plan.training
        dst.needs_grad = dst.needs_grad if gi else False
        dst.needs_grad = dst.needs_grad or self.needs_grad

    def _off(self, plan):
        return self.gi * self.Bg * self.H * self.W * self.dst.ld * plan.prec.esize

    def emit_fwd(self, plan):
        pr = plan.prec
        d, off = self.dst, self._off(plan)

        def run(st, self=self):
            src = plan.cur_inputs[self.slot]
            N.call.vae2_nchw_to_act(src.data_ptr(), d.ptr + off, pr.code, self.Bg, self.C, d.Cp, self.H, self.W, d.ld,
                                    self.src_ctot, self.src_coff, st)
        plan.pre_fwd.append(run)

    def emit_bwd(self, plan):
        if not self.needs_grad:
            return
        pr = plan.prec
        d, off = self.dst, self._off(plan)

        def run(st, self=self):
            dst = plan.cur_input_grads[self.slot]      # zero-filled by the autograd node; windows / groups accumulate
            N.call.vae2_act_to_nchw(d.grad().ptr + off, dst.data_ptr(), pr.code, self.Bg, self.C, self.H, self.W, d.ld,
                                    self.src_ctot, self.src_coff, 1, st)
        plan.post_bwd.append(run)


class CodeOp:
    """Per-sample code [B,Z,1,1] broadcast over H x W into a slice (enc_hrnet.py:454-462). Eager."""

    def reads(self):
        return []

    def writes(self):
        return [self.dst]

    def __init__(self, plan, slot, Z, dst, needs_grad=False):
        self.slot, self.Z, self.dst = slot, Z, dst
        self.needs_grad = needs_grad and plan.training       # a per-sample z that is a function of the posterior net
        dst.needs_grad = self.needs_grad

    def emit_fwd(self, plan):
        pr = plan.prec

        def run(st, self=self):
            code = plan.cur_inputs[self.slot]
            d = self.dst
            N.call.vae2_code_broadcast(code.data_ptr(), d.ptr, pr.code, d.B, self.Z, d.Cp, d.H, d.W, d.ld, st)
        plan.pre_fwd.append(run)

    def emit_bwd(self, plan):
        if not self.needs_grad:
            return
        pr, d = plan.prec, self.dst

        def run(st, self=self):      # d z[b][c] += sum over the pixels of the map the code was repeated into
            dst = plan.cur_input_grads[self.slot]
            N.call.vae2_spatial_sum(d.grad().ptr, dst.data_ptr(), pr.code, 1, d.B, d.H * d.W, self.Z, d.ld, self.Z, 1.0, 1, st)
        plan.post_bwd.append(run)


class OutputOp:
    """Channels-last activation -> channel window of an external NCHW fp32 tensor.  Eager."""

    def reads(self):
        return [self.act]

    def writes(self):
        return []

    def __init__(self, plan, slot, act, dst_ctot, dst_coff):
        self.slot, self.act, self.dst_ctot, self.dst_coff = slot, act, dst_ctot, dst_coff

    def emit_fwd(self, plan):
        pr = plan.prec

        def run(st, self=self):
            a = self.act
            dst = plan.cur_outputs[self.slot]
            N.call.vae2_act_to_nchw(a.ptr, dst.data_ptr(), pr.code, a.B, a.C, a.H, a.W, a.ld, self.dst_ctot,
                                    self.dst_coff, 0, st)
        plan.post_fwd.append(run)

    def emit_bwd(self, plan):
        pr = plan.prec
        a = self.act
        g = a.grad()
        a.take_acc_flag()   # the upstream gradient is the first writer of this buffer

        def run(st, self=self):
            src = plan.cur_output_grads[self.slot]
            if src is None:           # output unused by the loss: its gradient is zero
                g.buf.zero_()
                return
            N.call.vae2_nchw_to_act(src.data_ptr(), g.ptr, pr.code, a.B, a.C, a.Cp, a.H, a.W, a.ld, self.dst_ctot,
                                    self.dst_coff, st)
        plan.pre_bwd.append(run)


class ConvOp:
    """nn.Conv2d 3x3 (s1/s2, p1) or 1x1, optional bias."""

    def reads(self):
        return [self.x]

    def writes(self):
        return [self.y]

    def __init__(self, plan, x, conv, y=None):
        k, s = conv.kernel_size[0], conv.stride[0]
        self.conv, self.x, self.k, self.s, self.taps = conv, x, k, s, k * k
        Ho, Wo = conv_out(x.H, k, s), conv_out(x.W, k, s)
        self.y = y if y is not None else plan.new_act(conv.out_channels, Ho, Wo, name="conv", B=x.B)
        assert self.y.B == x.B
        assert (self.y.H, self.y.W) == (Ho, Wo)
        assert conv.in_channels == x.C, (conv.in_channels, x.C)
        plan.param(conv.weight)
        if conv.bias is not None:
            plan.param(conv.bias)
        self.w_off = 0
        self.engine = 0

    def _geom(self):
        x, y = self.x, self.y
        g = N.ConvGeom(B=x.B, H=x.H, W=x.W, Cin_p=x.Cp, ldx=x.ld, Ho=y.H, Wo=y.W, Cout_p=y.Cp, ldy=y.ld,
                       k=self.k, stride=self.s, pad=self.k // 2)
        g.Cin, g.Cout = self.conv.in_channels, self.conv.out_channels   # logical widths (host-side only: profiler)
        return g

    def emit_fwd(self, plan):
        pr = plan.prec
        g = self._geom()
        plan.keep.append(g)
        x, y = self.x, self.y
        wp = (plan.wq_flat.data_ptr() + 2 * self.w_off) if self.engine == 1 else (plan.wp_flat.data_ptr() + 4 * self.w_off)
        if self.engine == 2:
            wp = plan.w3f_flat.data_ptr() + 2 * self.x3_foff
        bias = None
        if self.conv.bias is not None:
            # bias padded to Cout_p lanes (pad = 0); refreshed from the parameter every forward
            bp = torch.zeros(y.Cp, dtype=torch.float32, device=plan.device)
            plan.keep.append(bp)
            b = self.conv.bias
            n = b.numel()
            plan.fwd.append(lambda st: bp[:n].copy_(b.detach()))
            bias = bp.data_ptr()
        gp = C.byref(g)
        xp, yp, eng = x.ptr, y.ptr, self.engine
        plan.fwd.append(lambda st: N.call.vae2_conv2d_fwd(xp, wp, bias, yp, pr.code, gp, eng, st))

    def emit_bwd(self, plan):
        pr = plan.prec
        g = self._geom()
        plan.keep.append(g)
        gp = C.byref(g)
        x, y = self.x, self.y
        dy = y.grad()
        dyp, xp = dy.ptr, x.ptr
        dwp = plan.dwp_flat.data_ptr() + 4 * self.w_off
        if not self.conv.weight.requires_grad:
            pass          # frozen parameter (e.g. a discriminator inside the generator step): no weight gradient
        elif getattr(self, "wgrad_tc", False):
            wsp = plan.wgrad_ws.data_ptr()
            plan.bwd.append(lambda st: N.call.vae2_conv2d_wgrad_tc(xp, dyp, dwp, wsp, gp, st))
        elif getattr(self, "wgrad_x2", False):
            wsp = plan.wgrad_x2_ws.data_ptr()
            plan.bwd.append(lambda st: N.call.vae2_conv2d_wgrad_f32x2(xp, dyp, dwp, wsp, gp, st))
        else:
            plan.bwd.append(lambda st: N.call.vae2_conv2d_wgrad(xp, dyp, dwp, pr.code, gp, 0, st))
        if self.conv.bias is not None and self.conv.bias.requires_grad:
            db = plan.grad_ptr(self.conv.bias)
            npix, cb, ldy = y.npix, y.C, y.ld
            plan.bwd.append(lambda st: N.call.vae2_bias_grad(dyp, db, pr.code, npix, cb, ldy, 0, st))
        if x.needs_grad:
            eng = self.engine
            wpT = (plan.wqT_flat.data_ptr() + 2 * self.w_off) if eng == 1 else (plan.wpT_flat.data_ptr() + 4 * self.w_off)
            if eng == 2:
                wpT = plan.w3b_flat.data_ptr() + 2 * self.x3_boff
            acc = x.take_acc_flag()
            dxp = x.grad().ptr
            plan.bwd.append(lambda st: N.call.vae2_conv2d_dgrad(dyp, wpT, dxp, pr.code, gp, acc, eng, st))


class _BnPart:
    """One statistics group of a BN: a contiguous range of samples normalised with its own batch statistics."""
    pass


class BnOp:
    """BatchNorm2d (+residual)(+ReLU) on a raw conv output; training or eval statistics.

    Emission is split into sub-steps so that a BnGroupOp can run the members of a SyncBN group through
    ONE collective: stats(+rank-local merge) | all-gather | finalize+apply, and in backward
    reduce | all-reduce | coeffs+elemt.  Without a collective (single rank, batch statistics) each direction
    is ONE cooperative launch (vae2_bn_fwd_fused / vae2_bn_bwd_fused).

    Statistic groups (plan.stat_groups = G > 1): the batch holds G stacked calls of the reference (e.g. the three
    per-frame discriminator passes, lib/utils/utils.py:116-119) -- samples [g*B/G, (g+1)*B/G) form group g, which is
    normalised with its OWN batch statistics; running statistics take G momentum updates in group order and
    num_batches_tracked advances by G, exactly as G sequential module calls do; d(gamma), d(beta) sum over groups."""

    _ws = {}

    def reads(self):
        return [self.y] + ([self.res] if self.res is not None else [])

    def writes(self):
        return [self.out]

    def __init__(self, plan, y, bn, relu, residual=None, out=None):
        self.y, self.bn, self.relu, self.res = y, bn, relu, residual
        self.out = out if out is not None else plan.new_act(y.C, y.H, y.W, name="bn", B=y.B)
        assert self.out.Cp == y.Cp or out is not None
        plan.param(bn.weight)
        plan.param(bn.bias)
        self.sync = isinstance(bn, torch.nn.SyncBatchNorm) and plan.world_size > 1

    # ---- forward sub-steps ---------------------------------------------------------------------
    def _fwd_setup(self, plan):
        y, out, res = self.y, self.out, self.res
        f32 = dict(dtype=torch.float32, device=plan.device)
        self.batch_stats = plan.bn_batch_stats or not self.bn.track_running_stats
        G = plan.stat_groups if self.batch_stats else 1
        assert y.B % G == 0
        self.npix_g = y.npix // G
        es = plan.prec.esize
        stat = torch.zeros(G, 6, y.Cp, **f32)              # per group: mean, invstd, scale, shift, c1, c2
        plan.keep.append(stat)
        self.parts = []
        for gi in range(G):
            pt = _BnPart()
            off = gi * self.npix_g
            pt.yp, pt.outp = y.ptr + off * y.ld * es, out.ptr + off * out.ld * es
            pt.resp = (res.ptr + off * res.ld * es) if res is not None else None
            pt.off, pt.first = off, gi == 0
            pt.mean, pt.invstd, pt.scale, pt.shift, pt.c1, pt.c2 = (stat[gi, j].data_ptr() for j in range(6))
            self.parts.append(pt)
        mode = os.environ.get("VAE2_BN_SPLIT", "0")          # "1": split both directions, "fwd" / "bwd": one of them
        fuse_max = int(os.environ.get("VAE2_BN_FUSE_MAX_MB", "1000000")) << 20
        nbytes = self.npix_g * y.Cp * (4 if plan.prec.code == 0 else 2)
        can = self.batch_stats and not self.sync and self.out.Cp >= y.Cp and nbytes <= fuse_max
        self.fused = can and mode not in ("1", "fwd")
        self.fused_bwd = can and mode not in ("1", "bwd")
        if self.batch_stats and not self.fused:
            for pt in self.parts:
                pt.partials = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * y.Cp, **f32)
                pt.npart = C.c_int(0)

    @staticmethod
    def _scratch(plan):
        """Partials scratch of the fused kernels: consumed inside the launch that writes it, so one buffer
        per device serves every BN of every plan."""
        key = plan.device.index
        ws = BnOp._ws.get(key)
        if ws is None:
            ws = BnOp._ws[key] = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * 2048, dtype=torch.float32,
                                             device=plan.device)
        return ws

    def _bn_ptrs(self):
        bn = self.bn
        rm = bn.running_mean.data_ptr() if bn.running_mean is not None else None
        rv = bn.running_var.data_ptr() if bn.running_var is not None else None
        nbt = bn.num_batches_tracked.data_ptr() if bn.num_batches_tracked is not None else None
        mom = 0.0 if bn.momentum is None else float(bn.momentum)
        return bn.weight.data_ptr(), bn.bias.data_ptr(), rm, rv, nbt, mom, float(bn.eps)

    def _emit_fwd_fused(self, plan):
        """stats | finalize | apply as ONE cooperative launch per statistics group (single-rank training statistics)."""
        pr, y, out, res = plan.prec, self.y, self.out, self.res
        ws = self._scratch(plan).data_ptr()
        ldr = res.ld if res is not None else 0
        npix, ldy, ldo, C_, Cp = self.npix_g, y.ld, out.ld, y.C, y.Cp
        gp, bp, rm, rv, nbt, mom, eps = self._bn_ptrs()
        relu = 1 if self.relu else 0
        p0, G = self.parts[0], len(self.parts)
        if G > 1 and os.environ.get("VAE2_BN_GROUP_LAUNCH", "1") != "0":
            # all statistics groups in ONE cooperative launch (csrc/bn.cu: the grid is split evenly between the groups)
            plan.fwd.append(lambda st: N.call.vae2_bn_fwd_fused_groups(
                p0.yp, p0.resp, p0.outp, ws, pr.code, npix, C_, Cp, ldy, ldr, ldo, gp, bp, rm, rv, nbt, mom, eps,
                p0.mean, p0.invstd, p0.scale, p0.shift, relu, G, 6 * Cp, st))
            return
        for pt in self.parts:
            plan.fwd.append(lambda st, pt=pt: N.call.vae2_bn_fwd_fused(
                pt.yp, pt.resp, pt.outp, ws, pr.code, npix, C_, Cp, ldy, ldr, ldo, gp, bp, rm, rv, nbt, mom, eps,
                pt.mean, pt.invstd, pt.scale, pt.shift, relu, st))

    def _emit_bwd_fused(self, plan):
        pr = plan.prec
        y, out, res, g = self.y, self.out, self.res, self.g
        npix, lanes, C_ = self.npix_g, self.lanes, y.C
        ws = self._scratch(plan).data_ptr()
        want_p = self.param_grads
        dgam = plan.grad_ptr(self.bn.weight) if want_p else None
        dbet = plan.grad_ptr(self.bn.bias) if want_p else None
        dy = y.grad()
        acc_dy = y.take_acc_flag()
        has_dres, ld_dres, acc_res = False, 0, 0
        if res is not None and res.needs_grad:
            acc_res = res.take_acc_flag()
            has_dres, ld_dres = True, res.ld
        # ReLU mask: without a residual it is recomputed from y (mode 2) and the stored activation is not read
        relu = 0 if not self.relu else (1 if res is not None else 2)
        gld, old, yld, dld = g.ld, out.ld, y.ld, dy.ld
        es = pr.esize
        p0, G = self.parts[0], len(self.parts)
        if G > 1 and os.environ.get("VAE2_BN_GROUP_LAUNCH", "1") != "0":
            dres0 = res.grad().ptr if has_dres else None
            gp0, dyp0 = g.ptr, dy.ptr
            plan.bwd.append(lambda st: N.call.vae2_bn_bwd_fused_groups(
                gp0, p0.outp, p0.yp, dyp0, dres0, ws, pr.code, npix, C_, lanes, gld, old, yld, dld, ld_dres, p0.mean,
                p0.invstd, p0.scale, p0.shift, dgam, dbet, 0, p0.c1, p0.c2, relu, acc_dy, acc_res, G, 6 * y.Cp, st))
            return
        for pt in reversed(self.parts):
            gp_, dyp = g.ptr + pt.off * gld * es, dy.ptr + pt.off * dld * es
            dres_p = (res.grad().ptr + pt.off * ld_dres * es) if has_dres else None
            accp = 0 if pt is self.parts[-1] else 1          # groups run last-to-first: the last group writes, the rest add
            plan.bwd.append(lambda st, pt=pt, gp_=gp_, dyp=dyp, dres_p=dres_p, accp=accp: N.call.vae2_bn_bwd_fused(
                gp_, pt.outp, pt.yp, dyp, dres_p, ws, pr.code, npix, C_, lanes, gld, old, yld, dld, ld_dres, pt.mean,
                pt.invstd, pt.scale, pt.shift, dgam, dbet, accp, pt.c1, pt.c2, relu, acc_dy, acc_res, st))

    def _emit_stats(self, plan, merged_ptrs=None):
        """Per-CTA partials (and, for SyncBN, the rank-local merge into the group message)."""
        pr, y = plan.prec, self.y
        Cp, npix, ld = y.Cp, self.npix_g, y.ld
        for i, pt in enumerate(self.parts):
            pp = pt.partials.data_ptr()
            plan.fwd.append(lambda st, pt=pt, pp=pp: N.call.vae2_bn_stats(pt.yp, pp, C.byref(pt.npart), pr.code, npix, Cp, ld, st))
            if merged_ptrs is not None:
                mp = merged_ptrs[i]
                plan.fwd.append(lambda st, pt=pt, pp=pp, mp=mp: N.call.vae2_bn_merge(pp, pt.npart.value, Cp, mp, st))

    def _emit_finalize(self, plan, parts_ptrs=None, n_parts=None, stride=0):
        y = self.y
        Cp, C_ = y.Cp, y.C
        gp, bp, rm, rv, nbt, mom, eps = self._bn_ptrs()
        if self.batch_stats:
            for i, pt in enumerate(self.parts):      # group order = the reference's call order (running-stat updates)
                outs = (pt.mean, pt.invstd, pt.scale, pt.shift)
                if parts_ptrs is None:
                    pp = pt.partials.data_ptr()
                    plan.fwd.append(lambda st, pt=pt, pp=pp, outs=outs: N.call.vae2_bn_finalize(
                        pp, pt.npart.value, C_, Cp, gp, bp, rm, rv, nbt, mom, eps, *outs, st))
                else:
                    src = parts_ptrs[i]
                    plan.fwd.append(lambda st, src=src, outs=outs: N.call.vae2_bn_finalize_strided(
                        src, n_parts, stride, C_, Cp, gp, bp, rm, rv, nbt, mom, eps, *outs, st))
        else:
            pt = self.parts[0]
            plan.fwd.append(lambda st: N.call.vae2_bn_eval_coeffs(C_, Cp, gp, bp, rm, rv, eps, pt.scale, pt.shift, st))

    def _emit_apply(self, plan):
        pr = plan.prec
        y, out, res = self.y, self.out, self.res
        lanes = min(y.Cp, out.Cp)
        ldr = res.ld if res is not None else 0
        relu = 1 if self.relu else 0
        npix, ldy, ldo = (self.npix_g if self.batch_stats else y.npix), y.ld, out.ld
        for pt in self.parts:
            plan.fwd.append(lambda st, pt=pt: N.call.vae2_bn_apply(pt.yp, pt.resp, pt.outp, pr.code, npix, lanes, ldy, ldr,
                                                                  ldo, pt.scale, pt.shift, relu, st))

    def _emit_fwd_peer(self, plan):
        """SyncBN forward as ONE cooperative launch: the cross-rank merge runs inside it over NVLink peer memory
        (engine/peer.py, csrc/bn.cu phase 3) -- no collective call."""
        pr, y, out, res = plan.prec, self.y, self.out, self.res
        ws = self._scratch(plan).data_ptr()
        ldr = res.ld if res is not None else 0
        gp, bp, rm, rv, nbt, mom, eps = self._bn_ptrs()
        relu = 1 if self.relu else 0
        p0, G, npix = self.parts[0], len(self.parts), self.npix_g
        slot, seq = peer.alloc(G, y.Cp, False)
        plan.fwd.append(lambda st: N.call.vae2_bn_fwd_fused_peer(
            p0.yp, p0.resp, p0.outp, ws, pr.code, npix, y.C, y.Cp, y.ld, ldr, out.ld, gp, bp, rm, rv, nbt, mom, eps, p0.mean,
            p0.invstd, p0.scale, p0.shift, relu, G, 6 * y.Cp, slot, seq, st))

    def _emit_bwd_peer(self, plan):
        pr = plan.prec
        y, out, res, g = self.y, self.out, self.res, self.g
        npix, lanes, C_ = self.npix_g, self.lanes, y.C
        ws = self._scratch(plan).data_ptr()
        want_p = self.param_grads
        dgam = plan.grad_ptr(self.bn.weight) if want_p else None
        dbet = plan.grad_ptr(self.bn.bias) if want_p else None
        dy = y.grad()
        acc_dy = y.take_acc_flag()
        has_dres, ld_dres, acc_res = False, 0, 0
        if res is not None and res.needs_grad:
            acc_res = res.take_acc_flag()
            has_dres, ld_dres = True, res.ld
        relu = 0 if not self.relu else (1 if res is not None else 2)
        p0, G = self.parts[0], len(self.parts)
        dres0 = res.grad().ptr if has_dres else None
        gp0, dyp0 = g.ptr, dy.ptr
        inv_count = 1.0 / (npix * plan.world_size)
        slot, seq = (0, 0) if plan._gsim else peer.alloc(G, y.Cp, True)      # the sizing pass of the backward emits nothing real
        plan.bwd.append(lambda st: N.call.vae2_bn_bwd_fused_peer(
            gp0, p0.outp, p0.yp, dyp0, dres0, ws, pr.code, npix, C_, lanes, g.ld, out.ld, y.ld, dy.ld, ld_dres, p0.mean,
            p0.invstd, p0.scale, p0.shift, dgam, dbet, 0, p0.c1, p0.c2, relu, acc_dy, acc_res, G, 6 * y.Cp, inv_count,
            slot, seq, st))

    def _emit_sync_fwd1(self, plan, msg_ptr):
        """SyncBN forward, first half: statistics + rank-local merge of every group in ONE cooperative launch."""
        pr, y = plan.prec, self.y
        ws = self._scratch(plan).data_ptr()
        p0, G, npix = self.parts[0], len(self.parts), self.npix_g
        plan.fwd.append(lambda st: N.call.vae2_bn_sync_fwd_stats(p0.yp, ws, pr.code, npix, y.C, y.Cp, y.ld, G, msg_ptr, st))

    def _emit_sync_fwd2(self, plan, gathered_ptr, world, stride):
        """second half: finalize every group from the `world` gathered sets (+ running statistics) and apply."""
        pr, y, out, res = plan.prec, self.y, self.out, self.res
        ldr = res.ld if res is not None else 0
        gp, bp, rm, rv, nbt, mom, eps = self._bn_ptrs()
        relu = 1 if self.relu else 0
        p0, G, npix = self.parts[0], len(self.parts), self.npix_g
        plan.fwd.append(lambda st: N.call.vae2_bn_sync_fwd_apply(
            p0.yp, p0.resp, p0.outp, pr.code, npix, y.C, y.Cp, y.ld, ldr, out.ld, gp, bp, rm, rv, nbt, mom, eps, p0.mean,
            p0.invstd, p0.scale, p0.shift, relu, G, 6 * y.Cp, gathered_ptr, world, stride, st))

    def _emit_sync_bwd(self, plan, phase, msg_ptr):
        """SyncBN backward halves (phase 1 before, phase 2 after the all-reduce), every group in one cooperative launch."""
        pr = plan.prec
        y, out, res, g = self.y, self.out, self.res, self.g
        npix, lanes, C_ = self.npix_g, self.lanes, y.C
        ws = self._scratch(plan).data_ptr()
        want_p = self.param_grads
        dgam = plan.grad_ptr(self.bn.weight) if want_p else None
        dbet = plan.grad_ptr(self.bn.bias) if want_p else None
        relu = 0 if not self.relu else (1 if res is not None else 2)
        p0, G = self.parts[0], len(self.parts)
        has_dres = res is not None and res.needs_grad
        if phase == 1:
            # (the elementwise outputs are not touched in this half; the residual-gradient pointer only selects the variant)
            self._sync_state = (y.grad(), res.grad() if has_dres else None)
            dyb, dres = self._sync_state
            gp0, dyp0, dres0 = g.ptr, dyb.ptr, (dres.ptr if dres is not None else None)
            plan.bwd.append(lambda st: N.call.vae2_bn_sync_bwd(
                1, gp0, p0.outp, p0.yp, dyp0, dres0, ws, pr.code, npix, C_, lanes, g.ld, out.ld, y.ld, dyb.ld,
                res.ld if has_dres else 0, p0.mean, p0.invstd, p0.scale, p0.shift, dgam, dbet, 0, p0.c1, p0.c2, relu, 0, 0, G,
                6 * y.Cp, msg_ptr, None, 0.0, st))
            return
        dyb, dres = self._sync_state
        acc_dy = y.take_acc_flag()
        acc_res = res.take_acc_flag() if has_dres else 0
        gp0, dyp0, dres0 = g.ptr, dyb.ptr, (dres.ptr if dres is not None else None)
        inv_count = 1.0 / (npix * plan.world_size)
        plan.bwd.append(lambda st: N.call.vae2_bn_sync_bwd(
            2, gp0, p0.outp, p0.yp, dyp0, dres0, ws, pr.code, npix, C_, lanes, g.ld, out.ld, y.ld, dyb.ld,
            res.ld if has_dres else 0, p0.mean, p0.invstd, p0.scale, p0.shift, None, None, 0, p0.c1, p0.c2, relu, acc_dy,
            acc_res, G, 6 * y.Cp, None, msg_ptr, inv_count, st))

    def emit_fwd(self, plan):
        """Stand-alone BN (its own collective under SyncBN)."""
        BnGroupOp(plan, [self]).emit_fwd(plan)

    # ---- backward sub-steps --------------------------------------------------------------------
    def _bwd_setup(self, plan):
        y = self.y
        f32 = dict(dtype=torch.float32, device=plan.device)
        # d(gamma), d(beta) are skipped when neither is trainable (e.g. the discriminators inside the generator step)
        self.param_grads = self.bn.weight.requires_grad or self.bn.bias.requires_grad
        for pt in self.parts:
            if not self.fused_bwd:
                pt.bparts = torch.zeros(N.lib().vae2_bn_max_partials() * 2 * y.Cp, **f32)
            pt.sums = torch.zeros(2 * y.Cp, **f32)
            pt.bnpart = C.c_int(0)
        self.lanes = min(y.Cp, self.out.Cp)
        self.g = self.out.grad()

    def _emit_bwd_reduce(self, plan):
        pr = plan.prec
        y, out, g = self.y, self.out, self.g
        npix, lanes, C_ = self.npix_g, self.lanes, y.C
        relu = 1 if self.relu else 0
        gld, old, yld = g.ld, out.ld, y.ld
        es = pr.esize
        for pt in self.parts:
            gp_ = g.ptr + pt.off * gld * es
            pp, sp = pt.bparts.data_ptr(), pt.sums.data_ptr()
            plan.bwd.append(lambda st, pt=pt, gp_=gp_, pp=pp: N.call.vae2_bn_bwd_reduce(
                gp_, pt.outp, pt.yp, pp, C.byref(pt.bnpart), pr.code, npix, lanes, gld, old, yld, pt.mean, pt.invstd, relu, st))
            plan.bwd.append(lambda st, pt=pt, pp=pp, sp=sp: N.call.vae2_bn_bwd_finalize(pp, pt.bnpart.value, C_, lanes, sp, st))

    def _emit_bwd_apply(self, plan, gsum_ptrs=None):
        """coefficients (from the global sums) + elementwise pass; parameter grads from the LOCAL sums."""
        pr = plan.prec
        y, out, res, g = self.y, self.out, self.res, self.g
        npix, lanes, C_ = self.npix_g, self.lanes, y.C
        want_p = self.param_grads
        dgam = plan.grad_ptr(self.bn.weight) if want_p else None
        dbet = plan.grad_ptr(self.bn.bias) if want_p else None
        count = npix * (plan.world_size if gsum_ptrs is not None else 1)
        dy = y.grad()
        acc_dy = y.take_acc_flag()
        has_dres, ld_dres, acc_res = False, 0, 0
        if res is not None and res.needs_grad:
            acc_res = res.take_acc_flag()
            has_dres, ld_dres = True, res.ld
        relu = 1 if self.relu else 0
        gld, old, yld, dld = g.ld, out.ld, y.ld, dy.ld
        es = pr.esize
        for i, pt in enumerate(self.parts):
            sp = pt.sums.data_ptr()
            gs = gsum_ptrs[i] if gsum_ptrs is not None else sp
            accp = 0 if i == 0 else 1
            plan.bwd.append(lambda st, pt=pt, gs=gs, sp=sp, accp=accp: N.call.vae2_bn_bwd_coeffs(
                gs, C_, lanes, 1.0 / count, dgam, dbet, accp, sp, pt.c1, pt.c2, st))
            gp_, dyp = g.ptr + pt.off * gld * es, dy.ptr + pt.off * dld * es
            dres_p = (res.grad().ptr + pt.off * ld_dres * es) if has_dres else None
            plan.bwd.append(lambda st, pt=pt, gp_=gp_, dyp=dyp, dres_p=dres_p: N.call.vae2_bn_bwd_elemt(
                gp_, pt.outp, pt.yp, dyp, dres_p, pr.code, npix, lanes, gld, old, yld, dld, ld_dres, pt.mean, pt.invstd,
                pt.scale, pt.c1, pt.c2, relu, acc_dy, acc_res, st))

    def emit_bwd(self, plan):
        BnGroupOp(plan, [self]).emit_bwd(plan)


class BnGroupOp:
    """BNs at the same depth of independent branches.  Single rank: the members simply run one after the
    other.  SyncBN: their (count, mean, M2) messages are concatenated and travel in ONE all-gather, their
    backward sums in ONE all-reduce -- the reference's per-BN collectives (1177 + 1160 per iteration,
    SURVEY.md §2.2) shrink by the branch count (and by the statistics-group count of a stacked plan)."""

    def reads(self):
        return [a for m in self.members for a in m.reads()]

    def writes(self):
        return [a for m in self.members for a in m.writes()]

    def __init__(self, plan, members):
        self.members = members
        self.sync = bool(members) and members[0].sync

    def emit_fwd(self, plan):
        ms = self.members
        for m in ms:
            m._fwd_setup(plan)
        if not (self.sync and ms[0].batch_stats):
            for m in ms:
                if m.fused:
                    m._emit_fwd_fused(plan)
                    continue
                if m.batch_stats:
                    m._emit_stats(plan)
                m._emit_finalize(plan)
                m._emit_apply(plan)
            return
        if peer.active() and all(m.out.Cp >= m.y.Cp for m in ms):
            for m in ms:                     # exchange inside each launch: no collective, nothing to group
                m._emit_fwd_peer(plan)
            plan.n_peer_bn_fwd = getattr(plan, "n_peer_bn_fwd", 0) + len(ms)
            return
        f32 = dict(dtype=torch.float32, device=plan.device)
        offs, total = [], 0
        for m in ms:
            o = []
            for _ in m.parts:
                o.append(total)
                total += 3 * m.y.Cp
            offs.append(o)
        msg, gathered = torch.zeros(total, **f32), torch.zeros(plan.world_size * total, **f32)
        plan.keep += [msg, gathered]
        # two cooperative launches per BN (all statistics groups in each) around the collective instead of
        # stats | merge | finalize | apply per group (VAE2_SYNCBN_HALVES=0 restores the split kernels)
        halves = os.environ.get("VAE2_SYNCBN_HALVES", "1") != "0" and all(m.out.Cp >= m.y.Cp for m in ms)
        for m, o in zip(ms, offs):
            if halves:
                m._emit_sync_fwd1(plan, msg.data_ptr() + 4 * o[0])
            else:
                m._emit_stats(plan, merged_ptrs=[msg.data_ptr() + 4 * x for x in o])
        grp = plan.group
        plan.fwd.append(lambda st: dist.all_gather_into_tensor(gathered, msg, group=grp))
        plan.n_collectives_fwd += 1
        for m, o in zip(ms, offs):
            if halves:
                m._emit_sync_fwd2(plan, gathered.data_ptr() + 4 * o[0], plan.world_size, total)
            else:
                m._emit_finalize(plan, parts_ptrs=[gathered.data_ptr() + 4 * x for x in o], n_parts=plan.world_size, stride=total)
                m._emit_apply(plan)

    def emit_bwd(self, plan):
        ms = list(reversed(self.members))
        for m in ms:
            m._bwd_setup(plan)
        if not (self.sync and ms[0].batch_stats):
            for m in ms:
                if m.fused_bwd:
                    m._emit_bwd_fused(plan)
                    continue
                m._emit_bwd_reduce(plan)
                m._emit_bwd_apply(plan)
            return
        if peer.active() and all(m.lanes == m.y.Cp for m in ms):
            for m in ms:
                m._emit_bwd_peer(plan)
            if not plan._gsim:
                plan.n_peer_bn_bwd = getattr(plan, "n_peer_bn_bwd", 0) + len(ms)
            return
        f32 = dict(dtype=torch.float32, device=plan.device)
        offs, total = [], 0
        for m in ms:
            o = []
            for _ in m.parts:
                o.append(total)
                total += 2 * m.y.Cp
            offs.append(o)
        gsum = torch.zeros(total, **f32)
        plan.keep.append(gsum)
        halves = os.environ.get("VAE2_SYNCBN_HALVES", "1") != "0" and all(m.lanes == m.y.Cp for m in ms)
        for m, o in zip(ms, offs):
            if halves:
                m._emit_sync_bwd(plan, 1, gsum.data_ptr() + 4 * o[0])
                continue
            m._emit_bwd_reduce(plan)
            for pt, off in zip(m.parts, o):
                dst, src, n = gsum[off:off + 2 * m.lanes], pt.sums, 2 * m.lanes
                plan.bwd.append(lambda st, dst=dst, src=src, n=n: dst.copy_(src[:n]))
        grp = plan.group
        plan.bwd.append(lambda st: dist.all_reduce(gsum, group=grp))
        plan.n_collectives_bwd += 1
        for m, o in zip(ms, offs):
            if halves:
                m._emit_sync_bwd(plan, 2, gsum.data_ptr() + 4 * o[0])
            else:
                m._emit_bwd_apply(plan, gsum_ptrs=[gsum.data_ptr() + 4 * x for x in o])


class FuseOp:
    """out = [relu](sum_j resize(src_j)); sources at the output size are added as they are,
    others are bilinearly resized (align_corners=False).  enc_hrnet.py:233-248 / :833-839."""

    def reads(self):
        return list(self.srcs)

    def writes(self):
        return [self.out]

    def __init__(self, plan, srcs, H, W, relu, out=None):
        self.srcs, self.relu = srcs, relu
        C_ = srcs[0].C
        self.out = out if out is not None else plan.new_act(C_, H, W, name="fuse", B=srcs[0].B)
        self.lanes = min([s.Cp for s in srcs] + [self.out.Cp])

    def emit_fwd(self, plan):
        pr = plan.prec
        arr = (N.FuseSrc * len(self.srcs))()
        for i, s in enumerate(self.srcs):
            arr[i] = N.FuseSrc(ptr=s.ptr, H=s.H, W=s.W, ld=s.ld)
        plan.keep.append(arr)
        o, n, relu, lanes = self.out, len(self.srcs), 1 if self.relu else 0, self.lanes
        op_ = o.ptr
        plan.fwd.append(lambda st: N.call.vae2_fuse_sum(arr, n, op_, pr.code, o.B, o.H, o.W, lanes, o.ld, relu, st))

    def emit_bwd(self, plan):
        pr = plan.prec
        o = self.out
        g = o.grad()
        relu, lanes = 1 if self.relu else 0, self.lanes
        same = [s for s in self.srcs if (s.H, s.W) == (o.H, o.W) and s.needs_grad]
        ups = [s for s in self.srcs if (s.H, s.W) != (o.H, o.W) and s.needs_grad]
        gp_, op_ = g.ptr, o.ptr
        if same:
            arr = (N.FuseDst * len(same))()
            for i, s in enumerate(same):
                acc = s.take_acc_flag()
                arr[i] = N.FuseDst(ptr=s.grad().ptr, ld=s.ld, accumulate=acc)
            plan.keep.append(arr)
            n, npix = len(same), o.npix
            plan.bwd.append(lambda st: N.call.vae2_fuse_bwd_same(gp_, op_, arr, n, pr.code, npix, lanes, g.ld, o.ld,
                                                                 relu, st))
        for s in ups:
            acc = s.take_acc_flag()
            sp = s.grad().ptr
            plan.bwd.append(lambda st, s=s, sp=sp, acc=acc: N.call.vae2_fuse_bwd_up(
                gp_, op_, sp, pr.code, o.B, o.H, o.W, s.H, s.W, lanes, g.ld, o.ld, s.ld, relu, acc, st))


class CopyOp:
    """dst slice (=) src   (channel concat done as a slice write)."""

    def reads(self):
        return [self.src]

    def writes(self):
        return [self.dst]

    def __init__(self, plan, src, dst):
        self.src, self.dst = src, dst

    def emit_fwd(self, plan):
        pr = plan.prec
        s, d = self.src, self.dst
        lanes = min(s.Cp, d.Cp)
        plan.fwd.append(lambda st: N.call.vae2_slice_copy(s.ptr, d.ptr, pr.code, s.npix, lanes, s.ld, d.ld, 0, st))

    def emit_bwd(self, plan):
        pr = plan.prec
        s, d = self.src, self.dst
        if not s.needs_grad:
            return
        lanes = min(s.Cp, d.Cp)
        acc = s.take_acc_flag()
        gs, gd = s.grad().ptr, d.grad().ptr
        plan.bwd.append(lambda st: N.call.vae2_slice_copy(gd, gs, pr.code, s.npix, lanes, d.ld, s.ld, acc, st))


class PoolOp:
    """nn.AdaptiveAvgPool2d((1, 1)): out[b][0][0][c] = mean over the pixels of x[b]  (non-HD_Z posterior head,
    enc_hrnet.py:1023-1041)."""

    def __init__(self, plan, x, out=None):
        self.x = x
        self.out = out if out is not None else plan.new_act(x.C, 1, 1, name="pool", B=x.B)

    def reads(self):
        return [self.x]

    def writes(self):
        return [self.out]

    def emit_fwd(self, plan):
        pr, x, o = plan.prec, self.x, self.out
        lanes, hw = min(x.Cp, o.Cp), x.H * x.W
        plan.fwd.append(lambda st: N.call.vae2_spatial_sum(x.ptr, o.ptr, pr.code, 0, x.B, hw, lanes, x.ld, o.ld, 1.0 / hw, 0, st))

    def emit_bwd(self, plan):
        pr, x, o = plan.prec, self.x, self.out
        if not x.needs_grad:
            return
        lanes, hw = min(x.Cp, o.Cp), x.H * x.W
        acc = x.take_acc_flag()
        gp, dxp = o.grad().ptr, x.grad().ptr
        plan.bwd.append(lambda st: N.call.vae2_spatial_bcast(gp, dxp, pr.code, 0, x.B, hw, lanes, x.Cp, x.ld, o.ld, 1.0 / hw, acc, st))


class TileOp:
    """dst[k*Bs + b] = src[b] for k < K: a feature computed once per context clip feeds all K latent draws stacked along
    the batch axis (K-sample inference, SURVEY.md §8 f1: the encoder trunk does not depend on z).  Forward only."""

    def reads(self):
        return [self.src]

    def writes(self):
        return [self.dst]

    def __init__(self, plan, src, dst, K):
        assert dst.B == K * src.B and (src.H, src.W) == (dst.H, dst.W)
        self.src, self.dst, self.K = src, dst, K
        dst.needs_grad = False

    def emit_fwd(self, plan):
        pr = plan.prec
        s, d = self.src, self.dst
        lanes = min(s.Cp, d.Cp)
        step = s.npix * d.ld * pr.esize
        for k in range(self.K):
            dp = d.ptr + k * step
            plan.fwd.append(lambda st, dp=dp: N.call.vae2_slice_copy(s.ptr, dp, pr.code, s.npix, lanes, s.ld, d.ld, 0, st))

    def emit_bwd(self, plan):
        if plan.training:
            raise RuntimeError("vae2_b200: TileOp (K-sample inference) has no backward")


# =============================================================================================
# recording front-end used by the nn.Module mirrors
# =============================================================================================
class Recorder:
    def __init__(self, plan):
        self.plan = plan
        self.n_in = 0
        self.n_out = 0

    # inputs / outputs ---------------------------------------------------------------------
    def input(self, C_, H, W, needs_grad, src_ctot=None, src_coff=0, slot=None, into=None, B=None):
        """Declare (a channel window of) external input #slot; `into` lists slice acts to fill."""
        p = self.plan
        if slot is None:
            slot = self.n_in
            self.n_in += 1
        dsts = into if into is not None else [p.new_act(C_, H, W, name="in%d" % slot, B=B)]
        p.add(InputOp(p, slot, C_, H, W, src_ctot if src_ctot is not None else C_, src_coff, dsts, needs_grad))
        return dsts[0] if into is None else dsts

    def group_input(self, slot, C_, H, W, src_ctot, src_coff, dst, gi, Bg, needs_grad):
        self.plan.add(GroupInputOp(self.plan, slot, C_, H, W, src_ctot, src_coff, dst, gi, Bg, needs_grad))

    def new_input_slot(self):
        s = self.n_in
        self.n_in += 1
        return s

    def code(self, slot, Z, dst, needs_grad=False):
        self.plan.add(CodeOp(self.plan, slot, Z, dst, needs_grad))

    def pool(self, x):
        return self.plan.add(PoolOp(self.plan, x)).out

    def output(self, act, dst_ctot=None, dst_coff=0, slot=None):
        if slot is None:
            slot = self.n_out
            self.n_out += 1
        self.plan.add(OutputOp(self.plan, slot, act, dst_ctot if dst_ctot is not None else act.C, dst_coff))
        return slot

    def new_output_slot(self):
        s = self.n_out
        self.n_out += 1
        return s

    # compute ------------------------------------------------------------------------------
    def conv(self, x, conv, y=None):
        return self.plan.add(ConvOp(self.plan, x, conv, y)).y

    def bn(self, y, bn, relu=False, residual=None, out=None):
        return self.plan.add(BnOp(self.plan, y, bn, relu, residual, out)).out

    def conv_bn(self, x, conv, bn, relu=False, residual=None, out=None):
        return self.bn(self.conv(x, conv), bn, relu, residual, out)

    def conv_bn_multi(self, items):
        """items: dicts {x, conv, bn, relu?, residual?, out?} for INDEPENDENT conv+BN pairs (same depth of
        different branches).  All convs are recorded first, then the BNs as one group (one SyncBN collective)."""
        if len(items) == 1:
            it = items[0]
            return [self.conv_bn(it["x"], it["conv"], it["bn"], it.get("relu", False), it.get("residual"), it.get("out"))]
        ys = [self.conv(it["x"], it["conv"]) for it in items]
        ops = [BnOp(self.plan, y, it["bn"], it.get("relu", False), it.get("residual"), it.get("out"))
               for y, it in zip(ys, items)]
        self.plan.add(BnGroupOp(self.plan, ops))
        return [op.out for op in ops]

    def fuse(self, srcs, H, W, relu=True, out=None):
        return self.plan.add(FuseOp(self.plan, srcs, H, W, relu, out)).out

    def copy(self, src, dst):
        self.plan.add(CopyOp(self.plan, src, dst))
        return dst

    def concat(self, Cs, H, W, name="cat", B=None):
        return self.plan.concat(Cs, H, W, name, B=B)

    def tile(self, src, dst, K):
        self.plan.add(TileOp(self.plan, src, dst, K))
        return dst
