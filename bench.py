#!/usr/bin/env python
"""bench.py -- VAE^2 training throughput on B200 (BASELINE.json metric: train frames/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host CPU cores

A step is one ``adversarial_train`` iteration of the reference (lib/core/function.py:491-512):
G-step (posterior net, reparam+KL, encoder + two decoders, discriminators, ELBO+GAN loss,
backward, Adam) followed by the D-step (2 + 6 discriminator passes, backward, Adam), on
synthetic Cityscapes-shaped clips.  One sample = (xt, x2t, x3t) = 9 RGB frames.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vae-2_b200", "lib")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (yaml, H, W, per-GPU batch, algorithmic conv FLOP per sample for a full iteration)
    # per-GPU batch: the largest that leaves ~30 GB of the 180 GB free (fp32 B=4 peaks at 148 GB, bf16 B=6 at 129 GB)
    "w18_256x512": ("vae2_hrnet_w18_small_v2_256x512.yaml", 256, 512, {"fp32": 4, "bf16": 6}, 7.60e12),
    "w18_1024x2048": ("vae2_hrnet_w18_small_v2_1024x2048.yaml", 1024, 2048, 1, 121.6e12),     # BASELINE configs[2]
    "w48_473x473": ("vae2_hrnet_w48_473x473.yaml", 473, 473, {"fp32": 1, "bf16": 2}, 126.9e12),   # configs[4], LIP
    "w48_520x520": ("vae2_hrnet_w48_520x520.yaml", 520, 520, {"fp32": 1, "bf16": 2}, 152.3e12),   # configs[4], PASCAL-Ctx
    "tiny_32x64": ("vae2_hrnet_tiny_32x64.yaml", 32, 64, 2, None),
}
FRAMES_PER_SAMPLE = 9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_models(cfg, dev, world, local_rank):
    import models.enc_hrnet as M
    import utils.utils as U
    import core.criterion as Cr
    nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
    g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss(),
                           cfg.TRAIN.X1RECON_LAMBDA, cfg.TRAIN.X2RECON_LAMBDA, cfg.TRAIN.X3RECON_LAMBDA,
                           cfg.TRAIN.GAN_LAMBDA)
    d = U.FullModel_D(nets[2], nets[3], Cr.lsgan_adversarial_loss())
    # The loop below is the reference's adversarial_train: optimizer_D.zero_grad() wipes whatever the generator step left
    # in the discriminators' .grad (function.py:499-512), so those gradients are not computed (config.skip_dead_D_grads)
    g.skip_dead_D_grads = True
    # trained-scale BN/conv init would need a checkpoint; the reference's own init is used (enc_hrnet.py:753-760)
    if world > 1:   # tools/train.py:216-229
        g = torch.nn.SyncBatchNorm.convert_sync_batchnorm(g)
        d = torch.nn.SyncBatchNorm.convert_sync_batchnorm(d)
    g, d = g.to(dev).train(), d.to(dev).train()
    gm, dm = g, d
    if world > 1:
        kw = dict(device_ids=[local_rank], output_device=local_rank, find_unused_parameters=True)
        # Two DDP switches on top of the reference's wrapping, neither changes a result: SyncBN running statistics are
        # bit-identical on every rank (tests/test_gpu_syncbn_peer.py), so re-broadcasting ~1000 buffers before every forward
        # is a no-op that costs 33 ms per iteration at N=2; bucket views drop one copy of every gradient (21 ms).
        kw.update(broadcast_buffers=False, gradient_as_bucket_view=True)
        for k_, v_ in (("broadcast_buffers", "VAE2_BENCH_DDP_BCAST"), ("gradient_as_bucket_view", "VAE2_BENCH_DDP_BUCKET_VIEW"),
                       ("static_graph", "VAE2_BENCH_DDP_STATIC")):
            if os.environ.get(v_) is not None:            # experiments only; the default is the reference's wrapping
                kw[k_] = os.environ[v_] == "1"
        d = torch.nn.parallel.DistributedDataParallel(d, **kw)
        g = torch.nn.parallel.DistributedDataParallel(g, **kw)
        from _engine_loader import engine
        if engine().peer.active():      # SyncBN exchanges inside the BN launches: keep NCCL kernels out of their way
            engine().peer.serialize_ddp(d)
            engine().peer.serialize_ddp(g)
    # tools/train.py:251-261: Adam, encdec optimizer excludes D parameters, D optimizer takes only them
    pg = [p for n, p in gm.named_parameters() if p.requires_grad and "D_model" not in n]
    pd = [p for n, p in dm.named_parameters() if p.requires_grad and "D_model" in n]
    opt_g = torch.optim.Adam(pg, lr=cfg.TRAIN.LR, fused=True)
    opt_d = torch.optim.Adam(pd, lr=cfg.TRAIN.LR, fused=True)
    return g, d, opt_g, opt_d


def reduce_tensor(t):
    """lib/core/function.py:32-43: the logging reduction of a loss to rank 0 (in place, under no_grad)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            dist.reduce(t, dst=0)
    return t


def train_step(g, d, opt_g, opt_d, xt, x2t, x3t):
    """One iteration of the reference's adversarial_train (lib/core/function.py:491-512)."""
    losses, _, x2p, _ = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0, is_baseline=False, baseline_mode="VAE_NATIVE")
    lg = reduce_tensor(losses[0])
    opt_g.zero_grad()
    losses[0].backward()
    opt_g.step()
    dl = d(x2t=x2t, x2t_predict=x2p.detach())
    ld = reduce_tensor(dl[0])
    opt_d.zero_grad()
    dl[0].backward()
    opt_d.step()
    return lg, ld


def step_roofline(E, dev, step_fn, hbm_peak, tf_peak, src):
    """Roofline of the kernel with the LARGEST TIME SHARE of the step (not a hand-picked shape): one extra iteration
    runs eagerly with every C-ABI launch bracketed by CUDA events on the launching stream (engine/profiler.py),
    grouped by kernel; achieved = that kernel's algorithmic FLOPs (bytes) summed over its launches / its summed time."""
    graphs = E.module._STATE["cuda_graphs"]
    E.use_cuda_graphs(False)
    try:
        step_fn()                                   # eager warm-up of the same plans (no re-recording)
        torch.cuda.synchronize(dev)
        with E.KernelProfile() as kp:
            step_fn()
    finally:
        E.use_cuda_graphs(graphs)
    fams = kp.by_kernel()
    total = sum(f["ms"] for f in fams) or 1.0
    top = fams[0]
    if os.environ.get("VAE2_BENCH_SHAPES"):         # per (kernel, shape) table of the profiled iteration, for tuning
        with open(os.environ["VAE2_BENCH_SHAPES"], "w") as fh:
            for r in kp.rows()[:120]:
                fh.write("%9.3f ms  n=%4d  %8.3f ms/launch  %7.1f GB/s %7.1f TF/s  %-58s %s\n" % (
                    r["ms"], r["n"], r["ms"] / r["n"], r["bytes"] / r["ms"] / 1e6, r["flop"] / r["ms"] / 1e9, r["kernel"], r["shape"]))

    def fr(r):
        t = r["ms"] * 1e-3
        tf, gb = r["flop"] / t / 1e12, r["bytes"] / t / 1e9
        t_tensor, t_hbm = r["flop"] / (tf_peak * 1e12), r["bytes"] / (hbm_peak * 1e9)
        return {"tflops": tf, "tensor_frac": tf / tf_peak, "gbs": gb, "hbm_frac": gb / hbm_peak,
                "bound": "tensor" if t_tensor >= t_hbm else "hbm", "frac": max(t_tensor, t_hbm) / t}
    f = fr(top)
    shp = top["top_shape"]
    fs = fr(shp)
    traffic, tsrc = ncu_traffic(top["kernel"], shp["shape"])
    roof = {"bound": f["bound"], "achieved": f["tflops"] if f["bound"] == "tensor" else f["gbs"],
            "peak": tf_peak if f["bound"] == "tensor" else hbm_peak, "unit": "TFLOP/s" if f["bound"] == "tensor" else "GB/s",
            "frac": f["frac"], "traffic": traffic, "traffic_source": tsrc,
            "kernel": top["kernel"], "share_of_step": top["ms"] / total, "launches": top["n"],
            "ms_per_launch": top["ms"] / top["n"], "tensor": {"achieved": f["tflops"], "peak": tf_peak, "frac": f["tensor_frac"]},
            "hbm": {"achieved": f["gbs"], "peak": hbm_peak, "frac": f["hbm_frac"]},
            "top_shape": {"shape": shp["shape"], "launches": shp["n"], "ms_per_launch": shp["ms"] / shp["n"],
                          "algorithmic_flop_per_launch": shp["flop"] / shp["n"],
                          "algorithmic_bytes_per_launch": shp["bytes"] / shp["n"], "tensor_frac": fs["tensor_frac"],
                          "hbm_frac": fs["hbm_frac"]},
            "peak_source": src + " (sustained bf16 / copy bandwidth; kernel timed inside the step)",
            "how": "CUDA events around every launch of one eager iteration, grouped by kernel; dominant by summed time"}
    table = [{"kernel": x["kernel"], "share": x["ms"] / total, "ms": x["ms"], "n": x["n"], **{k: v for k, v in fr(x).items()
              if k in ("tensor_frac", "hbm_frac", "bound", "frac")}} for x in fams[:12]]
    return roof, table, total


def ncu_traffic(kernel, shape):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full summary
    (profiles/ncu_traffic.json, written by tools/ncu_summarise.py from the .ncu-rep of the same build); None if that
    kernel has not been captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    try:
        db = json.load(open(path))
    except ValueError:
        return None, None
    base = kernel.split("+")[0].split("::")[-1].split(" ")[0]
    cands = [e for e in db.get("kernels", []) if base in e.get("kernel", "")]
    if not cands:
        return None, None
    # several captures may hold the kernel: take the newest one (tags sort by round letter: r2a < r2b < ...), and inside it the
    # launch whose label matches the shape, else its first launch
    newest = max(str(e.get("tag", "")) for e in cands)
    cands = [e for e in cands if str(e.get("tag", "")) == newest]
    best = next((e for e in cands if e.get("shape") == shape), cands[0])
    src = "profiles/%s_ncu_full.txt" % best["tag"] if best.get("tag") else db.get("source", "profiles/ncu_traffic.json")
    return best.get("dram_bytes"), "%s (%s, %s)" % (src, best.get("kernel"), best.get("shape"))


# ---- reference arm: the reference's own CPU implementation of the path on the box's host cores ------------------------
def _ref_tree():
    for p in (os.path.join(ROOT, "oracle", "_ref"), os.environ.get("VAE2_REFERENCE", "/root/reference")):
        if p and os.path.isfile(os.path.join(p, "lib", "models", "enc_hrnet.py")):
            return p
    return None


def cpu_reference_setup(cfg):
    """The UNMODIFIED reference modules (oracle/_ref, vendored by oracle/make_ref.py) when present -- kind
    'reference' -- else the oracle port driven by torch autograd -- kind 'port'.  Returns (kind, step_fn)."""
    import numpy as np
    ref = _ref_tree()
    Z = cfg.MODEL.EXTRA.Z_DIM
    if ref is not None:
        np.int = int                       # numpy >= 1.24 removed it (enc_hrnet.py:321,596,700)
        # the reference's lib/ must win over this repo's mirror for models / utils / core
        for m in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "core")]:
            del sys.modules[m]
        sys.path[:] = [p for p in sys.path if not p.endswith(os.path.join("vae-2_b200", "lib"))]
        sys.path.insert(0, os.path.join(ref, "lib"))
        import models.enc_hrnet as M
        import utils.utils as U
        import core.criterion as Cr
        assert ref in M.__file__, M.__file__
        nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
        g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss(),
                               cfg.TRAIN.X1RECON_LAMBDA, cfg.TRAIN.X2RECON_LAMBDA, cfg.TRAIN.X3RECON_LAMBDA,
                               cfg.TRAIN.GAN_LAMBDA).train()
        d = U.FullModel_D(nets[2], nets[3], Cr.lsgan_adversarial_loss()).train()
        og = torch.optim.Adam([p for n, p in g.named_parameters() if p.requires_grad and "D_model" not in n], lr=cfg.TRAIN.LR)
        od = torch.optim.Adam([p for n, p in d.named_parameters() if p.requires_grad and "D_model" in n], lr=cfg.TRAIN.LR)

        def step(xt, x2t, x3t):            # lib/core/function.py:491-512
            losses, _, x2p, _ = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0, is_baseline=False, baseline_mode="VAE_NATIVE")
            og.zero_grad()
            losses[0].backward()
            og.step()
            dl = d(x2t=x2t, x2t_predict=x2p.detach())
            od.zero_grad()
            dl[0].backward()
            od.step()
            return float(losses[0]) + float(dl[0])
        return "reference", step
    from oracle import vae2_oracle as O
    sd, opt = cpu_port_setup(cfg)

    def step(xt, x2t, x3t):
        B, _, H, W = xt.shape
        eps_z = [torch.randn(B, Z, h, w) for h, w in O.branch_sizes(H, W)]
        losses, _, x2p, _ = O.full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, torch.randn(B, Z, 1, 1))
        opt[0].zero_grad()
        losses[0].backward()
        opt[0].step()
        dl = O.full_d_forward(sd, cfg, x2t, x2p.detach())
        opt[1].zero_grad()
        dl[0].backward()
        opt[1].step()
        return float(losses[0]) + float(dl[0])
    return "port", step


def cpu_port_setup(cfg):
    import models.enc_hrnet as M
    import utils.utils as U
    import core.criterion as Cr
    nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
    g = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss())
    sd = {k: v.detach().clone() for k, v in g.state_dict().items()}   # reference-named state; modules not used further
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    pg = [v for k, v in sd.items() if v.requires_grad and "D_model" not in k]
    pd = [v for k, v in sd.items() if v.requires_grad and "D_model" in k]
    return sd, (torch.optim.Adam(pg, lr=1e-4), torch.optim.Adam(pd, lr=1e-4))


def cpu_baseline(cfg, H, W, steps=1, warmup=0, budget_s=420.0):
    """The reference's CPU implementation of the SAME step at the SAME frame size, B=1 (the bounded sample: one clip
    triple per step instead of the GPU arm's per-GPU batch; no size extrapolation), all host cores, fp32.  Steps are
    capped by a wall-clock budget so the run ends within minutes; the line says how many were timed."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, step = cpu_reference_setup(cfg)
    mk = lambda: (torch.randn(1, 9, H, W), torch.randn(1, 9, H, W), torch.randn(1, 9, H, W))
    t_start = time.time()
    for _ in range(warmup):
        step(*mk())
    times = []
    for _ in range(max(1, steps)):
        x = mk()
        t0 = time.time()
        step(*x)
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s:
            break
    dt = sorted(times)[len(times) // 2]
    what = ("the UNMODIFIED reference modules (oracle/_ref: FullModel_encdec + FullModel_D + Adam)" if kind == "reference"
            else "oracle port (functional restatement + torch autograd + Adam)")
    return {"value": FRAMES_PER_SAMPLE / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "%s, torch %s CPU fp32, %d thread(s): %d warm-up + %d timed full G+D iteration(s) at %dx%d, B=1 "
                      "(median %.1f s each); same frame size as the GPU arm, no scaling applied"
                      % (what, torch.__version__, cores, warmup, len(times), H, W, dt),
            "same_config": True, "steps_timed": len(times)}, dt


def reference_arm(args, cfg, H, W):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_baseline(cfg, H, W, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "train frames/sec", "value": base["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": base["steps_timed"], "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "net": NET_NAMES.get(args.workload, args.workload),
                       "step": "G-step + D-step (fwd, bwd, Adam)", "per_gpu_batch": 1, "global_batch": 1,
                       "frames_per_sample": FRAMES_PER_SAMPLE, "parallelism": "host cores (1 process)",
                       "steps_requested": args.steps, "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


NET_NAMES = {"w18_256x512": "VAE^2 HRNet-W18-small-v2 (encz + encdec + D_seq + D_frm)",
             "w18_1024x2048": "VAE^2 HRNet-W18-small-v2 (encz + encdec + D_seq + D_frm)",
             "w48_473x473": "VAE^2 HRNet-W48 (encz + encdec + D_seq + D_frm)",
             "w48_520x520": "VAE^2 HRNet-W48 (encz + encdec + D_seq + D_frm)",
             "tiny_32x64": "VAE^2 tiny HRNet (test config)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="w18_256x512", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bf16-path", action="store_true", help="skip the extra bf16 (tensor-core path) measurement")
    args = ap.parse_args()

    from config import load_config
    # discriminator calls stacked per pass: all of them at the small frame sizes, one at a time at 1024x2048 / W48 where a
    # stacked pass would not fit beside the generator networks (engine scratch region)
    os.environ.setdefault("VAE2_D_STACK", "1" if args.workload in ("w18_1024x2048", "w48_473x473", "w48_520x520") else "6")
    yaml_name, H, W, B, flop_per_sample = WORKLOADS[args.workload]
    if isinstance(B, dict):      # largest per-GPU batch that fits 180 GB with this precision's activation footprint
        B = B[args.precision]
    B = args.batch or B
    cfg = load_config(os.path.join(ROOT, "experiments", "vae2", yaml_name))
    if args.impl == "reference":
        return reference_arm(args, cfg, H, W)

    # The headline is BASELINE configs[1] (fp32).  The tensor-core (bf16) path is measured too, in a child
    # process that runs BEFORE this one allocates anything, and reported under "bf16_path" of the same line.
    bf16_path = None
    if (args.precision == "fp32" and not args.no_bf16_path and int(os.environ.get("WORLD_SIZE", "1")) == 1
            and args.workload == "w18_256x512"):
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--precision", "bf16", "--steps", str(args.steps),
                                  "--warmup", str(args.warmup), "--workload", args.workload, "--no-cpu-baseline"],
                                 capture_output=True, text=True, timeout=900)
            sub = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            bf16_path = {k: sub[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "roofline", "step_conv_tflops",
                                             "gpu_launches", "hbm_peak_gb") if k in sub}
            bf16_path["per_gpu_batch"] = sub["config"]["per_gpu_batch"]
        except Exception as e:   # the headline must not depend on the extra measurement
            bf16_path = {"unavailable": repr(e)[:200]}

    from _engine_loader import engine
    E = engine()
    E.native.lib()
    E.set_precision(args.precision)
    E.use_cuda_graphs(not args.no_graphs)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://")
        E.peer.enable()                 # SyncBN statistics over NVLink peer memory (VAE2_SYNCBN_P2P=0: NCCL collectives)
    torch.manual_seed(1234 + rank)
    g, d, opt_g, opt_d = build_models(cfg, dev, world, local_rank)

    # synthetic clips (N(0,1) frames, temporally correlated), several distinct batches resident in HBM
    nb = 2
    host = []
    for i in range(nb):
        xt = torch.randn(B, 9, H, W)
        x2t = xt + 0.1 * torch.randn(B, 9, H, W)
        x3t = x2t + 0.1 * torch.randn(B, 9, H, W)
        host.append(tuple(t.pin_memory() for t in (xt, x2t, x3t)))
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        train_step(g, d, opt_g, opt_d, *resident[i % nb])
    sync()
    l0 = E.launch_count()
    clocks = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu_range = os.environ.get("VAE2_BENCH_PROFILER_RANGE") == "1"   # `ncu --profile-from-start off`: the timed steps only
    if ncu_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        lg, ld = train_step(g, d, opt_g, opt_d, *resident[i % nb])
    e1.record()
    sync()
    if ncu_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = E.launch_count() - l0

    # end-to-end: host pinned inputs -> device every step, loss scalars read back every step
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        xb = tuple(t.to(dev, non_blocking=True) for t in host[i % nb])
        lg, ld = train_step(g, d, opt_g, opt_d, *xb)
        _ = (lg.item(), ld.item())
    f1.record()
    sync()
    ms_e2e = f0.elapsed_time(f1)
    clk = clocks.stop()
    E.check_finite(block=True)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    hbm, tf_burst, tf_sust, src = peaks()
    # every rank runs the profiled iteration (SyncBN / DDP collectives need all of them); rank 0 reports its own
    roof, kernel_table, prof_ms = step_roofline(E, dev, lambda: train_step(g, d, opt_g, opt_d, *resident[0]),
                                                hbm, tf_sust, src)
    if args.precision == "fp32" and roof["kernel"].startswith("t32::"):
        # fp32 products on bf16 tensor cores: every algorithmic FLOP is SIX issued bf16 FLOPs (exact 3-way split)
        # before lane padding; the tensor-pipe utilisation ncu measured for these kernels is in profiles/r2d_fp32_ncu_full.txt
        roof["tensor_issued"] = {"achieved": 6.0 * roof["tensor"]["achieved"], "peak": roof["tensor"]["peak"], "unit": "TFLOP/s",
                                 "frac": 6.0 * roof["tensor"]["frac"],
                                 "note": "6 bf16 products per fp32 product, lane padding not counted"}
    elif args.precision == "fp32":
        # exact-fp32 convs on the CUDA cores: the FMA-pipe fraction is the bound such a kernel can actually reach
        fma_peak = 148 * 128 * 2 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
        roof["fma_pipe"] = {"achieved": roof["tensor"]["achieved"], "peak": fma_peak, "unit": "TFLOP/s",
                            "frac": roof["tensor"]["achieved"] / fma_peak,
                            "peak_source": "148 SMs x 128 FP32 lanes x 2 x max SM clock"}
    if rank == 0:
        frames = FRAMES_PER_SAMPLE * B * world * args.steps
        value = frames / (ms * 1e-3)
        line = {
            "metric": "train frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "net": NET_NAMES.get(args.workload, args.workload),
                       "step": "G-step + D-step (fwd, bwd, Adam)", "per_gpu_batch": B, "global_batch": B * world,
                       "frames_per_sample": FRAMES_PER_SAMPLE, "parallelism": "dp%d" % world,
                       "l2": "activations per step >> 126 MB L2 (inputs larger than L2)",
                       "cuda_graphs": not args.no_graphs, "frame_size": [H, W], "skip_dead_D_grads": True,
                       "stacked_discriminator_passes": os.environ.get("VAE2_STACK_D", "1") != "0",
                       "d_stack": int(os.environ.get("VAE2_D_STACK", "6")),
                       "syncbn": ("peer-memory exchange inside the BN launches" if E.peer.active() else
                                  "NCCL all-gather / all-reduce per BN group") if world > 1 else "n/a",
                       "ddp": "find_unused_parameters=True (tools/train.py:226-229) + broadcast_buffers=False, "
                              "gradient_as_bucket_view=True" if world > 1 else "n/a"},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": 3 * B * 9 * H * W * 4, "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches),
            "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2),
            "arena_gb": E.ActArena.summary(),
            "clocks": clk,
            "roofline": roof,
            "kernel_shares": kernel_table,
        }
        if flop_per_sample:
            step_tf = flop_per_sample * B / (ms / args.steps * 1e-3) / 1e12
            line["step_conv_tflops"] = {"achieved": step_tf, "peak": tf_sust, "frac": step_tf / tf_sust,
                                        "note": "algorithmic conv FLOPs of the whole iteration / step time, per GPU; "
                                                "peak = sustained bf16 (" + src + ")"}
        if bf16_path is not None:
            line["bf16_path"] = bf16_path
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(cfg, H, W, steps=1, warmup=0)
        print(json.dumps(line), flush=True)
    if world > 1:
        E.peer.check()                  # raises if a BN launch gave up waiting for a peer rank
        # Tear down in order: drop the plans (their captured CUDA graphs hold the NCCL collectives), then the process
        # group.  Destroying a group whose collectives still sit inside live graphs was seen to hang at exit, so a
        # watchdog ends the process if the orderly path does not finish; every rank has passed the final barrier by now.
        dist.barrier()
        sys.stdout.flush()
        import gc
        import threading
        wd = threading.Timer(45.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        for m in list(g.modules()) + list(d.modules()):
            if hasattr(m, "reset_plans"):
                m.reset_plans()
        del g, d, opt_g, opt_d
        gc.collect()
        torch.cuda.synchronize(dev)
        E.peer.disable()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
