"""SyncBatchNorm exchange over NVLink peer memory (one node, one process per GPU).

torch's SyncBatchNorm (what `tools/train.py:217` converts every BN into) costs two NCCL collectives per layer and
iteration -- 1177 all-gathers and 1160 all-reduces of a few hundred bytes each for VAE^2.  With
`enable()` every rank maps every other rank's *mailbox* (one cudaMalloc per process, shared through CUDA IPC) and the
fused BN kernels of csrc/bn.cu exchange their per-channel partials with plain P2P stores INSIDE the launch: SyncBN then
is the same single cooperative launch per direction as BN on one GPU, and no collective is issued per layer.

Rules the host side has to keep (all checked or provided here):
  * the plans of all ranks must be recorded in the same order (they are: same model, same call sequence) -- slots in the
    mailbox and sequence counters are handed out by a bump allocator that must agree across ranks;
  * no NCCL kernel may be waiting for SMs while a BN launch waits for its peers (a cooperative launch owns the whole GPU):
    `serialize_ddp(ddp)` registers a DDP communication hook that makes the compute stream wait for each bucket's
    all-reduce before it goes on (the gradient all-reduce is ~0.15 ms per iteration here);
  * `check()` after a step raises if a launch gave up waiting for a peer (~60 s) instead of hanging the GPU.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import native as N

_MAX_OPS = 1 << 16
_S = {"ctx": None}


class _Ctx:
    pass


def active():
    return _S["ctx"] is not None


def world():
    return _S["ctx"].world if active() else 1


def enable(group=None, mailbox_mb=None):
    """Allocate + exchange the mailboxes.  Collective over `group` (default: WORLD); every rank must call it, on its own
    CUDA device, before any plan with SyncBatchNorm is recorded.  Returns False (and leaves the NCCL path in place) when the
    group has one rank, more than 8, or VAE2_SYNCBN_P2P=0."""
    if active():
        return True
    if not dist.is_available() or not dist.is_initialized() or os.environ.get("VAE2_SYNCBN_P2P", "1") == "0":
        return False
    W, rank = dist.get_world_size(group), dist.get_rank(group)
    if W < 2 or W > 8:
        return False
    lib = N.lib()
    nbytes = int(mailbox_mb if mailbox_mb is not None else os.environ.get("VAE2_PEER_MAILBOX_MB", "256")) << 20
    dev = torch.device("cuda", torch.cuda.current_device())
    ptr, handle = C.c_void_p(), C.create_string_buffer(64)
    ok = lib.vae2_ipc_alloc(nbytes, C.byref(ptr), handle) == 0
    mine = (os.uname().nodename, bytes(handle.raw) if ok else None)
    allh = [None] * W
    dist.all_gather_object(allh, mine, group=group)
    ok = ok and all(h[1] is not None and h[0] == mine[0] for h in allh)      # every rank allocated, all on this node
    ctx = _Ctx()
    ctx.world, ctx.rank, ctx.group, ctx.nbytes = W, rank, group, nbytes
    ctx.local, ctx.opened, bases = ptr, [], []
    for r in range(W):
        if not ok:
            break
        if r == rank:
            bases.append(ptr.value)
            continue
        p = C.c_void_p()
        if lib.vae2_ipc_open(C.create_string_buffer(allh[r][1], 64), C.byref(p)) != 0:
            ok = False                                 # no peer access / IPC not permitted here
            break
        ctx.opened.append(p)
        bases.append(p.value)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)                 # all ranks take the same path
    if int(flag.item()) == 0:
        for p in ctx.opened:
            lib.vae2_ipc_close(p)
        dist.barrier(group=group)
        if ptr.value:
            lib.vae2_ipc_free(ptr)
        return False                                   # SyncBN stays on NCCL collectives
    ctx.seq = torch.zeros(_MAX_OPS, dtype=torch.int32, device=dev)
    ctx.err = torch.zeros(1, dtype=torch.int32, device=dev)
    arr = (C.c_void_p * W)(*bases)
    N.check(lib.vae2_bn_peer_setup(W, rank, arr, ctx.seq.data_ptr(), ctx.err.data_ptr()), "vae2_bn_peer_setup")
    ctx.next_word, ctx.next_op = 0, 0
    torch.cuda.synchronize()
    dist.barrier(group=group)                           # every mailbox is zeroed and mapped before anyone stores into it
    _S["ctx"] = ctx
    return True


def alloc(groups, Cp, backward):
    """(slot_word, seq_index) of one BN op and direction; identical on every rank as long as plans are recorded in the
    same order."""
    ctx = _S["ctx"]
    words = N.lib().vae2_bn_peer_slot_words(ctx.world, groups, Cp, 1 if backward else 0)
    if (ctx.next_word + words) * 8 > ctx.nbytes or ctx.next_op >= _MAX_OPS:
        raise RuntimeError("vae2_b200: SyncBN peer mailbox exhausted (%d MB); raise VAE2_PEER_MAILBOX_MB" % (ctx.nbytes >> 20))
    out = (ctx.next_word, ctx.next_op)
    ctx.next_word += words
    ctx.next_op += 1
    return out


def check():
    """Raise if any BN launch of this rank gave up waiting for a peer (synchronizes the device)."""
    if active() and int(_S["ctx"].err.item()) != 0:
        raise RuntimeError("vae2_b200: a SyncBatchNorm launch timed out waiting for a peer rank (peer-memory exchange); "
                           "the ranks' plans diverged or a rank died")


def serialize_ddp(ddp, group=None):
    """DDP communication hook: the bucket all-reduce runs as usual, but the compute stream waits for it before launching
    anything else, so that no NCCL kernel is queued behind a BN launch that is itself waiting for the peer ranks."""
    W = dist.get_world_size(group)

    def hook(state, bucket):
        t = bucket.buffer()
        t.div_(W)
        dist.all_reduce(t, group=group, async_op=True).wait()
        fut = torch.futures.Future()
        fut.set_result(t)
        return fut
    ddp.register_comm_hook(None, hook)
    return ddp


def disable():
    ctx = _S["ctx"]
    if ctx is None:
        return
    torch.cuda.synchronize()
    lib = N.lib()
    lib.vae2_bn_peer_setup(0, 0, None, None, None)
    if dist.is_initialized():
        dist.barrier(group=ctx.group)
    for p in ctx.opened:
        lib.vae2_ipc_close(p)
    if dist.is_initialized():
        dist.barrier(group=ctx.group)
    lib.vae2_ipc_free(ctx.local)
    _S["ctx"] = None
