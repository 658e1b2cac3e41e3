"""GPU: SyncBatchNorm with the cross-rank exchange inside the launch, over peer memory (csrc/bn.cu phase 3,
engine/peer.py).  Two "ranks" are emulated on one GPU with one mailbox each; the real two-process run is
tests/dist_syncbn_check.py ... p2p (tests/test_gpu_dist.py)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from helpers import E, O, log_err, rel_err
from test_gpu_parity_full import _to_act, _from_act, _st, DEV

pytestmark = pytest.mark.gpu
N = E.native


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape,G,relu,with_res", [((8, 18, 17, 23), 2, True, True), ((4, 270, 9, 12), 1, False, False),
                                                   ((12, 64, 16, 32), 3, True, False)])
def test_syncbn_peer_memory_two_ranks_on_one_gpu(shape, G, relu, with_res, prec):
    """vae2_bn_fwd_fused_peer / vae2_bn_bwd_fused_peer.  Rank 0 runs first and finds rank 1's words already in its mailbox
    (pre-filled from the rank-local halves kernels -- the same merge code); rank 1 runs second and consumes the words
    RANK 0's LAUNCH stored into rank 1's mailbox, i.e. the real data path.  Must equal F.batch_norm over each group's
    whole batch (torch SyncBatchNorm, nn/modules/_functions.py:39-122), running statistics after G updates included;
    both ranks must end with bit-identical statistics; no launch may report a time-out."""
    code, tdt, al, tol = (0, torch.float32, 4, 2e-5) if prec == "fp32" else (1, torch.bfloat16, 8, 2e-2)
    B, C_, H, W = shape
    Bg, world = B // G, 2
    Bl = Bg // world
    Cp = (C_ + al - 1) // al * al
    tag = "sbp%s" % (shape,)
    y = O.det_normal(tag + "y", shape, 2.0, 0.5)
    y[::2] += 0.6
    res = O.det_normal(tag + "r", shape) if with_res else None
    go = O.det_normal(tag + "go", shape)
    if prec == "bf16":
        y, go = y.bfloat16().float(), go.bfloat16().float()
        res = res.bfloat16().float() if res is not None else None
    gam, bet = O.det_uniform(tag + "g", (C_,), 0.5, 1.5), O.det_normal(tag + "b", (C_,), 0.1)
    rm, rv = O.det_normal(tag + "rm", (C_,), 0.1), O.det_uniform(tag + "rv", (C_,), 0.5, 1.5)
    yr = y.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if res is not None else None
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    outs = []
    for gi in range(G):
        sl = slice(gi * Bg, (gi + 1) * Bg)
        o = F.batch_norm(yr[sl], rm_r, rv_r, gr, br, True, 0.01, 1e-5)
        o = o + rr[sl] if rr is not None else o
        outs.append(F.relu(o) if relu else o)
    ref = torch.cat(outs, 0)
    ref.backward(go)
    idx = [torch.cat([torch.arange(gi * Bg + r * Bl, gi * Bg + (r + 1) * Bl) for gi in range(G)]) for r in range(world)]
    f32 = dict(dtype=torch.float32, device=DEV)
    P = Bl * H * W
    ws = torch.zeros(N.lib().vae2_bn_max_partials() * 3 * Cp, **f32)
    ya = [_to_act(y[i], code, tdt, Cp) for i in idx]
    ga = [_to_act(go[i], code, tdt, Cp) for i in idx]
    ra = [_to_act(res[i], code, tdt, Cp) for i in idx] if res is not None else [None, None]
    lib = N.lib()
    wf, wb = lib.vae2_bn_peer_slot_words(world, G, Cp, 0), lib.vae2_bn_peer_slot_words(world, G, Cp, 1)
    assert wf == 2 * world * G * 3 * Cp and wb == 2 * world * G * 2 * Cp
    mbox = [torch.zeros(wf + wb, dtype=torch.int64, device=DEV) for _ in range(world)]
    seqs = [torch.zeros(2, dtype=torch.int32, device=DEV) for _ in range(world)]
    errs = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
    bases = (C.c_void_p * world)(*[m.data_ptr() for m in mbox])

    def as_rank(r):
        torch.cuda.synchronize()
        N.check(lib.vae2_bn_peer_setup(world, r, bases, seqs[r].data_ptr(), errs[r].data_ptr()), "vae2_bn_peer_setup")

    def prefill(dst_rank, src_rank, slot_word, nv, msg, seq):
        # word (buffer, src, group, value, lane) = {float bits, seq}; buffer = seq & 1
        base = slot_word + (((seq & 1) * world + src_rank) * G) * nv * Cp
        bits = msg.view(torch.int32).to(torch.int64) & 0xffffffff
        mbox[dst_rank][base:base + G * nv * Cp] = bits | (seq << 32)

    try:
        # ---- forward ----
        msg1 = torch.zeros(G * 3 * Cp, **f32)
        N.call.vae2_bn_sync_fwd_stats(ya[1].data_ptr(), ws.data_ptr(), code, P, C_, Cp, Cp, G, msg1.data_ptr(), _st())
        prefill(0, 1, 0, 3, msg1, 1)
        stat = [torch.zeros(G, 6, Cp, **f32) for _ in range(world)]
        run = [(rm.to(DEV), rv.to(DEV), torch.zeros(1, dtype=torch.int64, device=DEV)) for _ in range(world)]
        gd, bd = gam.to(DEV), bet.to(DEV)
        oa = [torch.zeros_like(t) for t in ya]
        for r in range(world):
            as_rank(r)
            sp = lambda j: stat[r][0, j].data_ptr()
            N.call.vae2_bn_fwd_fused_peer(ya[r].data_ptr(), ra[r].data_ptr() if ra[r] is not None else None, oa[r].data_ptr(),
                                          ws.data_ptr(), code, P, C_, Cp, Cp, Cp, Cp, gd.data_ptr(), bd.data_ptr(),
                                          run[r][0].data_ptr(), run[r][1].data_ptr(), run[r][2].data_ptr(), 0.01, 1e-5, sp(0),
                                          sp(1), sp(2), sp(3), 1 if relu else 0, G, 6 * Cp, 0, 0, _st())
        torch.cuda.synchronize()
        assert all(int(e) == 0 for e in errs), "a launch timed out waiting for its peer"
        assert all(int(s[0]) == 1 for s in seqs)
        out = torch.zeros(shape)
        for r in range(world):
            out[idx[r]] = _from_act(oa[r], code, len(idx[r]), C_, H, W, Cp)
        e_out = rel_err(out, ref.detach())
        e_run = max(rel_err(run[r][0].cpu(), rm_r) for r in range(world)) + max(rel_err(run[r][1].cpu(), rv_r) for r in range(world))
        assert all(int(run[r][2]) == G for r in range(world))
        assert torch.equal(stat[0][:, :4], stat[1][:, :4]), "ranks must merge to bit-identical statistics"
        # ---- backward ----
        mode = 0 if not relu else (1 if with_res else 2)
        dg = [torch.zeros(C_, **f32) for _ in range(world)]
        db = [torch.zeros(C_, **f32) for _ in range(world)]
        dya = [torch.zeros_like(t) for t in ya]
        dra = [torch.zeros_like(t) if res is not None else None for t in ya]
        gmsg1 = torch.zeros(G * 2 * Cp, **f32)
        sp1 = lambda j: stat[1][0, j].data_ptr()
        sdg, sdb = torch.zeros(C_, **f32), torch.zeros(C_, **f32)
        N.call.vae2_bn_sync_bwd(1, ga[1].data_ptr(), oa[1].data_ptr(), ya[1].data_ptr(), dya[1].data_ptr(),
                                dra[1].data_ptr() if dra[1] is not None else None, ws.data_ptr(), code, P, C_, Cp, Cp, Cp, Cp, Cp,
                                Cp, sp1(0), sp1(1), sp1(2), sp1(3), sdg.data_ptr(), sdb.data_ptr(), 0, sp1(4), sp1(5), mode, 0,
                                0, G, 6 * Cp, gmsg1.data_ptr(), None, 0.0, _st())
        prefill(0, 1, wf, 2, gmsg1, 1)
        for r in range(world):
            as_rank(r)
            sp = lambda j: stat[r][0, j].data_ptr()
            N.call.vae2_bn_bwd_fused_peer(ga[r].data_ptr(), oa[r].data_ptr(), ya[r].data_ptr(), dya[r].data_ptr(),
                                          dra[r].data_ptr() if dra[r] is not None else None, ws.data_ptr(), code, P, C_, Cp,
                                          Cp, Cp, Cp, Cp, Cp, sp(0), sp(1), sp(2), sp(3), dg[r].data_ptr(), db[r].data_ptr(),
                                          0, sp(4), sp(5), mode, 0, 0, G, 6 * Cp, 1.0 / (P * world), wf, 1, _st())
        torch.cuda.synchronize()
        assert all(int(e) == 0 for e in errs), "a launch timed out waiting for its peer"
        assert all(int(s[1]) == 1 for s in seqs)
    finally:
        torch.cuda.synchronize()
        lib.vae2_bn_peer_setup(0, 0, None, None, None)
    dx = torch.zeros(shape)
    dr = torch.zeros(shape)
    for r in range(world):
        dx[idx[r]] = _from_act(dya[r], code, len(idx[r]), C_, H, W, Cp)
        if res is not None:
            dr[idx[r]] = _from_act(dra[r], code, len(idx[r]), C_, H, W, Cp)
    e_dx = rel_err(dx, yr.grad)
    e_dg = rel_err(dg[0].cpu() + dg[1].cpu(), gr.grad)     # DDP sums / averages the per-rank parameter gradients
    e_db = rel_err(db[0].cpu() + db[1].cpu(), br.grad)
    e_dr = rel_err(dr, rr.grad) if res is not None else 0.0
    log_err("syncbn_peer_%s_%s_G%d" % (prec, "x".join(map(str, shape)), G), out=e_out, running=e_run, dx=e_dx, dgamma=e_dg,
            dbeta=e_db, dres=e_dr)
    assert e_out < tol and e_run < 2e-5, (e_out, e_run)
    assert e_dx < 5 * tol and e_dg < 5 * tol and e_db < 5 * tol and e_dr < 5 * tol, (e_dx, e_dg, e_db, e_dr)
