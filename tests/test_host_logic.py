"""CPU: the C-ABI library's export surface, the drop-in module surface, plan recording, config,
and the multi-rank host logic (world_size-2 gloo)."""
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import E, O, M, ROOT, build_product, cfg_of, golden


def test_c_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vae2_b200.h")).read()
    declared = set(re.findall(r"\b(vae2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 30
    h = E.native.lib()                       # loads libvae2_b200.so (no GPU needed to dlopen)
    missing = [n for n in sorted(declared) if not hasattr(h, n)]
    assert not missing, missing
    assert h.vae2_abi_version() == 1
    assert set(E.native.EXPORTS) <= declared


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(E.native, "_lib", None)
    monkeypatch.setattr(E.native, "LIB_PATH", "/nonexistent/libvae2_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        E.native.lib()


def test_cpu_tensors_are_rejected_not_emulated():
    cfg = cfg_of("vae2_hrnet_tiny_32x64.yaml")
    net = M.get_encz_model(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 18, 32, 64))
    import core.criterion as Cr
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Cr.L1Loss()(predict=torch.zeros(1, 3), target=torch.zeros(1, 3))


@pytest.mark.parametrize("name", ["tiny_b2_32x64", "w18_b1_32x64"])
def test_parameter_names_match_the_reference(name):
    gold = golden(name)
    g, d = build_product(cfg_of(str(gold["cfg"])))
    assert sorted(k for k, _ in g.named_parameters()) == sorted(gold["g_grad_names"].tolist())
    assert sorted(k for k, _ in d.named_parameters()) == sorted(gold["d_grad_names"].tolist())
    assert g.encz_model.hd_z and g.encz_model.z_dim == 8 and g.D_model_sequence.clip_length == 3


def test_plan_recording_covers_every_parameter():
    from vae2_b200_engine.graph import Plan, Recorder
    cfg = cfg_of("vae2_hrnet_w18_small_v2_256x512.yaml")
    net = M.get_encdec_model(cfg)
    plan = Plan(torch.device("cpu"), 1, "bf16", True, True)
    sizes = O.branch_sizes(64, 128)
    shapes = ((1, 9, 64, 128),) + tuple((1, 8, h, w) for h, w in sizes) + ((1, 8, 1, 1),)
    outs = net._record(Recorder(plan), shapes, (False, True, True, True, True, False), "")
    plan.finalize()
    assert outs == [(9, 64, 128)] * 3
    assert len(plan.params) == len(list(net.parameters()))
    kinds = [type(o).__name__ for o in plan.ops]
    n_bn = kinds.count("BnOp") + sum(len(o.members) for o in plan.ops if type(o).__name__ == "BnGroupOp")
    assert kinds.count("ConvOp") == 462 and n_bn == 453      # 462 convs / 453 BNs
    # BNs at the same depth of sibling branches share one group (= one SyncBN collective)
    assert kinds.count("BnOp") + kinds.count("BnGroupOp") < 0.6 * n_bn
    # concat lane maps: transition3_e sees [code 8 | z 8 | features C] in padded segments
    conv = next(o for o in plan.ops if type(o).__name__ == "ConvOp" and o.conv is net.transition3_e[0][0])
    assert conv.x.cin_map[:16] == list(range(16)) and conv.x.cin_map[16] == 16 and conv.x.Cp % 16 == 0


def test_config_surface(tmp_path):
    from config import get_cfg_defaults, update_config
    cfg = get_cfg_defaults()

    class A:
        cfg = os.path.join(ROOT, "experiments", "vae2", "vae2_hrnet_w18_small_v2_256x512.yaml")
        opts = ["TRAIN.LR", "0.001", "MODEL.EXTRA.Z_DIM", "8"]
    update_config(cfg, A)
    assert cfg.TRAIN.LR == 0.001 and cfg.MODEL.EXTRA["STAGE4"]["NUM_CHANNELS"] == [18, 36, 72, 144]
    assert cfg.MODEL.EXTRA.STAGE2.NUM_BRANCHES == 2 and cfg.DATASET.NUM_CLASSES == 3
    with pytest.raises(AttributeError):
        cfg.TRAIN.LR = 1.0                   # frozen
    with pytest.raises(KeyError):
        get_cfg_defaults().merge_from_list(["TRAIN.NO_SUCH_KEY", "1"])


# ---- world_size-2 gloo: the SyncBN message algebra and the data-parallel gradient semantics ----------
def _chan_merge(parts):
    """Python restatement of csrc/bn.cu chan_merge over a list of (count, mean, M2) rows."""
    n, mean, m2 = 0.0, 0.0, 0.0
    for nb, mb, m2b in parts:
        if nb <= 0:
            continue
        nn = n + nb
        d = mb - mean
        mean = mean + d * nb / nn
        m2 = m2 + m2b + d * d * n * nb / nn
        n = nn
    return n, mean, m2


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    C_, per = 5, 7 * 11
    full = torch.randn(world * per, C_, dtype=torch.float64) * 3 + 1.5      # the global batch, identical on all ranks
    mine = full[rank * per:(rank + 1) * per]
    # forward message: [3][C] = (count, mean, M2) per rank, all-gathered (graph.py BnOp.sync)
    msg = torch.stack([torch.full((C_,), float(per), dtype=torch.float64), mine.mean(0),
                       ((mine - mine.mean(0)) ** 2).sum(0)]).reshape(-1)
    gathered = torch.zeros(world * 3 * C_, dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, msg)
    parts = gathered.view(world, 3, C_)
    merged = [_chan_merge([(float(parts[r, 0, c]), float(parts[r, 1, c]), float(parts[r, 2, c])) for r in range(world)])
              for c in range(C_)]
    mean = torch.tensor([m[1] for m in merged], dtype=torch.float64)
    var = torch.tensor([m[2] / m[0] for m in merged], dtype=torch.float64)
    ok_fwd = torch.allclose(mean, full.mean(0), atol=1e-12) and torch.allclose(var, full.var(0, unbiased=False), atol=1e-12)
    # backward message: all-reduce of (sum dy, sum dy*xhat); dx uses the GLOBAL sums over the GLOBAL count
    dy_full = torch.randn(world * per, C_, dtype=torch.float64)
    xhat_full = (full - full.mean(0)) / torch.sqrt(full.var(0, unbiased=False) + 1e-5)
    dy, xhat = dy_full[rank * per:(rank + 1) * per], xhat_full[rank * per:(rank + 1) * per]
    sums = torch.stack([dy.sum(0), (dy * xhat).sum(0)])
    local = sums.clone()
    dist.all_reduce(sums)
    ok_bwd = torch.allclose(sums[0], dy_full.sum(0)) and torch.allclose(sums[1], (dy_full * xhat_full).sum(0))
    # parameter gradients stay LOCAL sums (DDP averages them afterwards): mean over ranks == global / world
    g = local[1].clone()
    dist.all_reduce(g)
    ok_ddp = torch.allclose(g / world, (dy_full * xhat_full).sum(0) / world)
    # bench.py's timing reduction: MAX over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, bool(ok_fwd), bool(ok_bwd), bool(ok_ddp), float(t)))
    dist.destroy_process_group()


def test_syncbn_message_algebra_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True, True, 11.0), (1, True, True, True, 11.0)]


def _ddp_hook_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2)).double()
    import copy
    a = torch.nn.parallel.DistributedDataParallel(copy.deepcopy(ref))
    b = torch.nn.parallel.DistributedDataParallel(copy.deepcopy(ref), broadcast_buffers=False, gradient_as_bucket_view=True)
    E.peer.serialize_ddp(b)                       # the hook the peer-memory SyncBN path installs (engine/peer.py)
    x = torch.randn(world, 4, 6, dtype=torch.float64)[rank]
    a(x).square().sum().backward()
    b(x).square().sum().backward()
    same = all(torch.allclose(pa.grad, pb.grad, atol=1e-14) for pa, pb in zip(a.parameters(), b.parameters()))
    q.put((rank, bool(same)))
    dist.destroy_process_group()


def test_serialized_ddp_hook_averages_like_default_ddp_world2_gloo():
    """engine.peer.serialize_ddp: same averaged gradients as DDP's built-in all-reduce (tools/train.py:226-229), only the
    stream ordering differs; also with the two constructor switches bench.py uses."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_ddp_hook_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_peer_mailbox_slot_words_and_inactive_default():
    """Host side of the peer-memory SyncBN exchange: slot sizes (2 buffers x world x groups x values x lanes 8-byte words)
    and that nothing is active unless enable() succeeded (single process: it must refuse)."""
    lib = E.native.lib()
    assert lib.vae2_bn_peer_slot_words(8, 6, 272, 0) == 2 * 8 * 6 * 3 * 272
    assert lib.vae2_bn_peer_slot_words(2, 1, 32, 1) == 2 * 2 * 1 * 2 * 32
    assert not E.peer.active() and E.peer.world() == 1
    assert E.peer.enable() is False               # no process group
    E.peer.check()                                # no-op when inactive


def test_activation_arena_shares_memory_across_phases_only():
    """engine.ActArena (host logic, CPU tensors): buffers of one phase are disjoint, the next phase walks the same
    chunks from the start, an allocation larger than a chunk gets a chunk of its own, and nothing is shared outside
    a phase or with VAE2_ACT_ARENA=0."""
    A = E.ActArena
    dev = torch.device("cpu")
    old_chunk = A.CHUNK_BYTES
    A.reset()
    A.CHUNK_BYTES = 4096 * 4          # 4096 floats per chunk
    try:
        assert A.alloc(dev, torch.float32, 100) is None, "no phase active: the caller allocates privately"

        def span(t):
            return (t.data_ptr(), t.data_ptr() + t.numel() * t.element_size())

        with E.activation_phase("G"):
            g = [A.alloc(dev, torch.float32, n) for n in (1000, 3000, 500, 3500, 700)]
        with E.activation_phase("D"):
            d = [A.alloc(dev, torch.float32, n) for n in (2000, 2000, 4000)]
        for group in (g, d):
            assert all(t.numel() == n for t, n in zip(group, [t.numel() for t in group]))
            assert all(float(t.abs().sum()) == 0.0 for t in group), "fresh buffers are zero"
            s = sorted(span(t) for t in group)
            assert all(a[1] <= b[0] for a, b in zip(s, s[1:])), "buffers of one phase must not overlap"
            bases = [c.data_ptr() for c in next(iter(A._pools.values()))["chunks"]]
            assert all(min((t.data_ptr() - b) % 512 for b in bases if b <= t.data_ptr()) == 0 for t in group), \
                "128-element alignment inside the chunk"
        assert d[0].data_ptr() == g[0].data_ptr(), "the second phase starts where the first one started"
        pool = next(iter(A._pools.values()))
        total = sum(c.numel() for c in pool["chunks"])
        assert total < sum(t.numel() for t in g + d), "footprint is below the sum over phases"
        with E.activation_phase("D"):
            big = A.alloc(dev, torch.float32, 10000)          # larger than a chunk: gets a chunk of its own
        assert big.numel() == 10000 and any(c.numel() >= 10000 for c in pool["chunks"])
        # nesting restores the outer phase; disabling turns sharing off
        with E.activation_phase("G"):
            with E.activation_phase("D"):
                pass
            assert A.phase == "G"
        assert A.phase is None
        os.environ["VAE2_ACT_ARENA"] = "0"
        with E.activation_phase("G"):
            assert A.alloc(dev, torch.float32, 10) is None
    finally:
        os.environ.pop("VAE2_ACT_ARENA", None)
        A.CHUNK_BYTES = old_chunk
        A.reset()


def test_activation_arena_recycles_dropped_plans():
    """A dropped plan's extents go back to its phase's free list: re-recording (new batch size, moved parameters,
    reset_plans) must not grow the arena (the bump allocator used to leak them)."""
    A = E.ActArena
    dev = torch.device("cpu")
    old_chunk = A.CHUNK_BYTES
    A.reset()
    A.CHUNK_BYTES = 8192 * 4

    class Owner:
        def __init__(self):
            self.arena_extents = []
    try:
        a, b = Owner(), Owner()
        with E.activation_phase("G"):
            ta = [A.alloc(dev, torch.float32, n, owner=a) for n in (1000, 2000)]
            tb = [A.alloc(dev, torch.float32, n, owner=b) for n in (500,)]
        pool = next(iter(A._pools.values()))
        before = sum(c.numel() for c in pool["chunks"])
        ptrs = sorted(t.data_ptr() for t in ta)
        A.give_back(a.arena_extents)
        with E.activation_phase("G"):
            c = Owner()
            tc = [A.alloc(dev, torch.float32, n, owner=c) for n in (2500, 400)]
        assert tc[0].data_ptr() == ptrs[0], "coalesced extents of the dropped plan are reused first-fit"
        assert sum(ch.numel() for ch in pool["chunks"]) == before, "no growth"
        spans = sorted((t.data_ptr(), t.data_ptr() + 4 * t.numel()) for t in tc + tb)
        assert all(x[1] <= y[0] for x, y in zip(spans, spans[1:])), "live buffers stay disjoint"
        with E.activation_phase("D"):
            td = A.alloc(dev, torch.float32, 100, owner=Owner())
        assert td.data_ptr() == pool["chunks"][0].data_ptr(), "other phases still start at the chunk base"
    finally:
        A.CHUNK_BYTES = old_chunk
        A.reset()
