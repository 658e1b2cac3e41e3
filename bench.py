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
    "w18_1024x2048": ("vae2_hrnet_w18_small_v2_1024x2048.yaml", 1024, 2048, 1, 121.6e12),
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
    # trained-scale BN/conv init would need a checkpoint; the reference's own init is used (enc_hrnet.py:753-760)
    if world > 1:   # tools/train.py:216-229
        g = torch.nn.SyncBatchNorm.convert_sync_batchnorm(g)
        d = torch.nn.SyncBatchNorm.convert_sync_batchnorm(d)
    g, d = g.to(dev).train(), d.to(dev).train()
    gm, dm = g, d
    if world > 1:
        d = torch.nn.parallel.DistributedDataParallel(d, device_ids=[local_rank], output_device=local_rank,
                                                      find_unused_parameters=True)
        g = torch.nn.parallel.DistributedDataParallel(g, device_ids=[local_rank], output_device=local_rank,
                                                      find_unused_parameters=True)
    # tools/train.py:251-261: Adam, encdec optimizer excludes D parameters, D optimizer takes only them
    pg = [p for n, p in gm.named_parameters() if p.requires_grad and "D_model" not in n]
    pd = [p for n, p in dm.named_parameters() if p.requires_grad and "D_model" in n]
    opt_g = torch.optim.Adam(pg, lr=cfg.TRAIN.LR, fused=True)
    opt_d = torch.optim.Adam(pd, lr=cfg.TRAIN.LR, fused=True)
    return g, d, opt_g, opt_d


def train_step(g, d, opt_g, opt_d, xt, x2t, x3t):
    losses, _, x2p, _ = g(xt=xt, x2t=x2t, x3t=x3t, multiplier=1.0, is_baseline=False, baseline_mode="VAE_NATIVE")
    opt_g.zero_grad()
    losses[0].backward()
    opt_g.step()
    dl = d(x2t=x2t, x2t_predict=x2p.detach())
    opt_d.zero_grad()
    dl[0].backward()
    opt_d.step()
    return losses[0], dl[0]


def conv_microbench(dev, E, steps=20):
    """The dominant launch of the step, timed alone with CUDA events on its own stream: the
    64->64 3x3 s1 conv at full resolution (16 % of the W18 MACs).  Returns
    (TFLOP/s algorithmic, ms per launch, description)."""
    import ctypes as C
    N = E.native
    prec = E.get_precision()
    code, tdt, al = (0, torch.float32, 4) if prec == "fp32" else (1, torch.bfloat16, 16)
    B, H, W, Cin, Cout, k = 1, 256, 512, 64, 64, 3
    x = torch.randn(B * H * W * Cin, device=dev).to(tdt)
    y = torch.zeros(B * H * W * Cout, dtype=tdt, device=dev)
    wp = torch.randn(k * k * Cin * Cout, device=dev) * 0.05
    eng = 0
    if prec == "bf16":   # tcgen05 path: packed bf16 weights [tap][Cout][Cin]
        wp, eng = wp.to(torch.bfloat16), 1
    g = N.ConvGeom(B=B, H=H, W=W, Cin_p=Cin, ldx=Cin, Ho=H, Wo=W, Cout_p=Cout, ldy=Cout, k=k, stride=1, pad=1)
    s = torch.cuda.Stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(s):
        for _ in range(3):
            N.call.vae2_conv2d_fwd(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), code, C.byref(g), eng, s.cuda_stream)
        for a, b in ev:
            flush.zero_()
            a.record(s)
            N.call.vae2_conv2d_fwd(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), code, C.byref(g), eng, s.cuda_stream)
            b.record(s)
    s.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[steps // 2]
    flop = 2.0 * B * H * W * Cin * Cout * k * k
    return flop / (ms * 1e-3) / 1e12, ms, "conv3x3 s1 64->64 @256x512 B=1 fwd (%s, %s, L2 flushed between launches)" % (prec, "tcgen05" if eng else "CUDA-core FMA")


def cpu_port_step(cfg, sd, opt, H, W, B, tag):
    """One G+D iteration of the reference algorithm on the host CPU (oracle port + torch autograd + Adam)."""
    from oracle import vae2_oracle as O
    Z = cfg.MODEL.EXTRA.Z_DIM
    xt, x2t, x3t = (torch.randn(B, 9, H, W) for _ in range(3))
    eps_z = [torch.randn(B, Z, h, w) for h, w in O.branch_sizes(H, W)]
    code = torch.randn(B, Z, 1, 1)
    losses, _, x2p, _ = O.full_encdec_forward(sd, cfg, xt, x2t, x3t, eps_z, code)
    opt[0].zero_grad()
    losses[0].backward()
    opt[0].step()
    dl = O.full_d_forward(sd, cfg, x2t, x2p.detach())
    opt[1].zero_grad()
    dl[0].backward()
    opt[1].step()
    return float(losses[0])


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


def cpu_baseline(cfg, H, W, steps=1, warmup=0):
    """Bounded sample: the same G+D iteration at (H/2, W/2), B=1; frames/s is scaled to the full
    frame size by the pixel ratio (conv cost is linear in pixels)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hs, ws = H // 2, W // 2
    sd, opt = cpu_port_setup(cfg)
    for _ in range(warmup):
        cpu_port_step(cfg, sd, opt, hs, ws, 1, "w")
    t0 = time.time()
    for _ in range(steps):
        cpu_port_step(cfg, sd, opt, hs, ws, 1, "t")
    dt = (time.time() - t0) / steps
    scale = (hs * ws) / float(H * W)
    return {"value": FRAMES_PER_SAMPLE / dt * scale, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "oracle port (torch %s CPU, fp32): %d full G+D iteration(s) at %dx%d B=1, %.1f s each; "
                      "frames/s scaled x%.2f to %dx%d frames" % (torch.__version__, steps, hs, ws, dt, scale, H, W)}, dt


def reference_arm(args, cfg, H, W):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_baseline(cfg, H, W, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "train frames/sec", "value": base["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "net": "VAE^2 HRNet-W18-small-v2 (encz + encdec + D_seq + D_frm)",
                       "step": "G-step + D-step (fwd, bwd, Adam)", "per_gpu_batch": 1, "global_batch": 1,
                       "frames_per_sample": FRAMES_PER_SAMPLE, "parallelism": "host cores (1 process)",
                       "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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
    e0.record()
    for i in range(args.steps):
        lg, ld = train_step(g, d, opt_g, opt_d, *resident[i % nb])
    e1.record()
    sync()
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
    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        tf, kms, desc = conv_microbench(dev, E)
        frames = FRAMES_PER_SAMPLE * B * world * args.steps
        value = frames / (ms * 1e-3)
        line = {
            "metric": "train frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "net": "VAE^2 HRNet-W18-small-v2 (encz + encdec + D_seq + D_frm)",
                       "step": "G-step + D-step (fwd, bwd, Adam)", "per_gpu_batch": B, "global_batch": B * world,
                       "frames_per_sample": FRAMES_PER_SAMPLE, "parallelism": "dp%d" % world,
                       "l2": "activations per step >> 126 MB L2 (inputs larger than L2); microbench flushes L2",
                       "cuda_graphs": not args.no_graphs},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": 3 * B * 9 * H * W * 4, "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches),
            "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2),
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_burst, "unit": "TFLOP/s", "frac": tf / tf_burst,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full capture
                         # summarised in profiles/r1_ncu_full_conv_tc_64x64.txt (bf16 kernel; output stays in L2)
                         "traffic": 16875520 if args.precision == "bf16" else None,
                         "kernel": desc, "ms_per_launch": kms, "peak_source": src + " (burst, kernel timed alone)"},
        }
        if args.precision == "fp32":
            # the exact-fp32 path runs on the CUDA cores: the tensor roofline above is the contract's yardstick,
            # the FMA-pipe one is the bound this kernel can actually reach
            fma_peak = 148 * 128 * 2 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
            line["roofline"]["fma_pipe"] = {"achieved": tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": tf / fma_peak,
                                            "peak_source": "148 SMs x 128 FP32 lanes x 2 x max SM clock"}
        if flop_per_sample:
            step_tf = flop_per_sample * B / (ms / args.steps * 1e-3) / 1e12
            line["step_conv_tflops"] = {"achieved": step_tf, "peak": tf_sust, "frac": step_tf / tf_sust,
                                        "note": "algorithmic conv FLOPs of the whole iteration / step time, per GPU; "
                                                "peak = sustained bf16 (" + src + ")"}
        if bf16_path is not None:
            line["bf16_path"] = bf16_path
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(cfg, H, W)
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: destroying a process group whose collectives live inside captured
        # CUDA graphs was seen to hang at exit; every rank has passed the final barrier by now
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
