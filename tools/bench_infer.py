#!/usr/bin/env python
"""BASELINE configs[3]: multi-sample prediction, K latent draws per context clip (tools/inference.py ->
lib/core/function.py::inference, :124-146 + metrics :238-316), 256x512, one B200.

    python tools/bench_infer.py [--precision fp32|bf16] [--K 16] [--clips 4] [--batch 1] [--reference-draws 2]

GPU arm: core.sampling.KSampleInference (one stacked pass per batch of clips: trunk once, K draws tiled, no posterior
net / discriminators, device-side L1 / PSNR / SSIM / MS-SSIM); inputs start in pinned host memory, the H2D copy and the
read-back of the per-draw scores are inside the timed region.  Reference arm: the unmodified reference wrapper in eval
mode, `prior_sampling`, on the host cores -- `--reference-draws` draws are timed and the per-clip time is that times
K / draws (the reference runs the K draws strictly one after another, so the scaling is exact; its CPU metrics, which
need pytorch_msssim, are NOT included in its time).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vae-2_b200", "lib"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--clips", type=int, default=6)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reference-draws", type=int, default=2)
    ap.add_argument("--size", default="256,512")
    args = ap.parse_args()
    H, W = (int(v) for v in args.size.split(","))
    from config import load_config
    from _engine_loader import engine
    import bench
    cfg = load_config(os.path.join(ROOT, "experiments", "vae2", "vae2_hrnet_w18_small_v2_256x512.yaml"))
    E = engine()
    E.native.lib()
    E.set_precision(args.precision)
    E.use_cuda_graphs(True)
    dev = torch.device("cuda", 0)
    g, d, _, _ = bench.build_models(cfg, dev, 1, 0)
    g.eval()
    import core.sampling as S
    drv = S.KSampleInference(g, K=args.K, with_ssim=True, keep_predictions=False)
    B = args.batch
    host = [tuple(torch.randn(B, 9, H, W).pin_memory() for _ in range(3)) for _ in range(2)]

    def one(i):
        xt, x2t, x3t = (t.to(dev, non_blocking=True) for t in host[i % 2])
        out = drv(xt, x2t, x3t)
        return float(out["x2t_psnr"].mean()) + float(out["x3t_ssim"].mean())     # device -> host read of the scores

    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    l0 = E.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.clips):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.clips
    launches = (E.launch_count() - l0) / args.clips
    draws_per_s = args.K * B / (ms * 1e-3)
    line = {"metric": "inference latent draws/sec (3 predicted clips = 9 frames per draw, with per-frame L1/PSNR/SSIM/MS-SSIM)",
            "value": draws_per_s, "unit": "draws/s", "frames_per_s": 9 * draws_per_s, "ms_per_clip_batch": ms, "K": args.K,
            "clips_per_batch": B, "dtype": "f32" if args.precision == "fp32" else "bf16", "n_gpus": 1, "data": "synthetic",
            "config": {"workload": "infer_k%d_%dx%d" % (args.K, H, W), "net": "VAE^2 HRNet-W18-small-v2 encdec (eval-mode BN)"},
            "gpu_launches_per_clip_batch": launches, "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2),
            "e2e": {"h2d_bytes_per_clip_batch": 3 * B * 9 * H * W * 4, "d2h_bytes_per_clip_batch": 16}}
    if args.reference_draws > 0:
        import numpy as np
        ref = bench._ref_tree()
        if ref is not None:
            np.int = int
            for m in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "core")]:
                del sys.modules[m]
            sys.path[:] = [p for p in sys.path if not p.endswith(os.path.join("vae-2_b200", "lib"))]
            sys.path.insert(0, os.path.join(ref, "lib"))
            import models.enc_hrnet as M
            import utils.utils as U
            import core.criterion as Cr
            torch.set_num_threads(os.cpu_count() or 1)
            nets = [M.get_encz_model(cfg), M.get_encdec_model(cfg), M.get_D_sequence_model(cfg), M.get_D_frame_model(cfg)]
            rg = U.FullModel_encdec(nets[0], nets[1], nets[2], nets[3], Cr.L1Loss(), Cr.KLLoss(), Cr.lsgan_adversarial_loss()).eval()
            xs = [torch.randn(1, 9, H, W) for _ in range(3)]
            with torch.no_grad():
                rg(xt=xs[0], x2t=xs[1], x3t=xs[2], multiplier=1.0, sampling_mode="prior_sampling")     # warm-up draw
                t0 = time.time()
                for _ in range(args.reference_draws):
                    rg(xt=xs[0], x2t=xs[1], x3t=xs[2], multiplier=1.0, sampling_mode="prior_sampling")
                dt = (time.time() - t0) / args.reference_draws
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "draws/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "unmodified reference FullModel_encdec.eval(), prior_sampling, %d timed draw(s) of "
                                              "one clip at %dx%d, %.2f s per draw; a K=%d clip costs K times that (the reference "
                                              "loops over draws); its CPU image metrics are not included" %
                                              (args.reference_draws, H, W, dt, args.K)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
