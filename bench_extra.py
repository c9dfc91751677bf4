#!/usr/bin/env python
"""Secondary measurements (NOT the driver's bench contract -- that is bench.py): the widened rows of
SURVEY.md section 8 timed on one GPU with CUDA events, inputs resident in HBM, one JSON line per workload:

  dh        depth-hints objective (A18, BASELINE config 4): B=32, 1024x320, [0,'s'], 4 scales, fwd+bwd
  md_f2     multi-source photometric objective (BASELINE config 5): B=16, [0,-1,1], 1024x320 and 2048x640
  costvol   ManyDepth cost volume (next-3): B=16, 2 lookups, 96 bins, 16 ch at 80x256 (1024x320 / 4)

usage: python bench_extra.py [--steps K] [--warmup W] [--only dh,md_f2,costvol]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="dh,md_f2,costvol")
    args = ap.parse_args()
    from depthmodelhardening_b200 import _lib, synth
    lib = _lib.load()
    dev = torch.device("cuda:0")
    peak = peak_gbs()
    want = args.only.split(",")

    if "dh" in want:
        from depthmodelhardening_b200 import depth_hints as DH
        B, H, W = 32, 320, 1024
        pb = synth.photo_batch(batch=B, height=H, width=W, frame_ids=(0, "s"), seed=5, depth_hints=True).to(dev)
        noise = {s: pb.noise[s][:, :1].contiguous() for s in pb.scales}
        disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}

        def step():
            for d in disps.values():
                d.grad = None
            losses, _ = DH.depth_hint_losses(pb.color, disps, pb.K, pb.inv_K, pb.T, pb.frame_ids, pb.scales, H, W,
                                             pb.extras["depth_hint"], pb.extras["depth_hint_mask"], True, noise=noise)
            losses["loss"].backward()
        n0 = lib.dmh_launch_count()
        ms = timed(step, args.steps, args.warmup)
        launches = (lib.dmh_launch_count() - n0) // (args.steps + args.warmup)
        # compulsory bytes / px: config-2 objective (152) + hint depth, valid, hint loss read per scale (12 x 4)
        bpp = 152.0 + 48.0
        print(json.dumps({"workload": "depth-hints objective fwd+bwd (config 4)", "B": B, "H": H, "W": W,
                          "ms_per_step": ms, "mpix_per_s": B * H * W / ms / 1e3, "gpu_launches": int(launches),
                          "algorithmic_bytes_per_px": bpp, "hbm_frac": bpp * B * H * W / (ms * 1e-3) / 1e9 / peak,
                          "note": "fused: one dmh_photo_scale_dh launch per scale (multi-source kernel in depth-hints mode)"}))

    if "md_f2" in want:
        from depthmodelhardening_b200 import objective
        for (B, H, W) in ((16, 320, 1024), (4, 640, 2048)):
            pb = synth.photo_batch(batch=B, height=H, width=W, frame_ids=(0, -1, 1), seed=6).to(dev)
            disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
            Ts = {k: v.clone().requires_grad_(True) for k, v in pb.T.items()}

            def step():
                for d in disps.values():
                    d.grad = None
                for t in Ts.values():
                    t.grad = None
                losses, _ = objective.photometric_losses(pb.color, disps, pb.K, pb.inv_K, Ts, pb.frame_ids, pb.scales,
                                                         H, W, noise=pb.noise)
                losses["loss"].backward()
            ms = timed(step, args.steps, args.warmup)
            bpp = 208.0          # F=2
            print(json.dumps({"workload": "photometric objective fwd+bwd, two temporal sources + pose gradients "
                                          "(config 5)", "B": B, "H": H, "W": W, "ms_per_step": ms,
                              "mpix_per_s": B * H * W / ms / 1e3, "algorithmic_bytes_per_px": bpp,
                              "hbm_frac": bpp * B * H * W / (ms * 1e-3) / 1e9 / peak,
                              "note": "photo_scale_kernel<2> (general multi-source kernel)"}))
            del pb, disps, Ts

    if "costvol" in want:
        from depthmodelhardening_b200 import cost_volume as CV
        cost_volume_inputs = synth.cost_volume_inputs
        B, L, h, w, D = 16, 2, 80, 256, 96
        cur, look, poses, K, inv_K, bins = [t.to(dev) for t in cost_volume_inputs(B=B, L=L, h=h, w=w, D=D, seed=9)]
        ms = timed(lambda: CV.cost_volume(cur, look, poses, K, inv_K, bins), args.steps, args.warmup)
        alg = B * h * w * ((1 + L) * 16 * 4 + 2 * D * 4)
        print(json.dumps({"workload": "ManyDepth cost volume (next-3)", "B": B, "lookups": L, "bins": D, "h": h, "w": w,
                          "ms_per_step": ms, "cells_per_s": B * D * h * w / (ms * 1e-3),
                          "algorithmic_bytes": alg, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                          "taps_gbs": B * D * L * h * w * 4 * 64 / (ms * 1e-3) / 1e9,
                          "note": "gather-bound: taps_gbs = bytes requested from L1/L2 by the bilinear taps"}))


if __name__ == "__main__":
    main()
