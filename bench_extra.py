#!/usr/bin/env python
"""Secondary measurements (NOT the driver's bench contract -- that is bench.py): the widened rows of
SURVEY.md section 8 timed on one GPU with CUDA events, inputs resident in HBM, one JSON line per workload:

  cfg1      BASELINE configs[0] (B=4, 640x192, [0,'s']): eager vs CUDA-graph replay (objective.GraphedObjective)
  dh        depth-hints objective (A18, BASELINE config 4): B=32, 1024x320, [0,'s'], 4 scales, fwd+bwd
  md_f2     multi-source photometric objective (BASELINE config 5): B=16, [0,-1,1], 1024x320 and 2048x640
  costvol   ManyDepth cost volume (next-3): B=16, 2 lookups, 96 bins, 16 ch at 80x256 (1024x320 / 4)
  attack    CONTEXT: the whole L0 / L-inf attack loop of the drop-in classes with a stock
            PyTorch random-init ResNet-18 monodepth2-style depth network in the loop (the network is outside
            the graft), Ba=32 scenes of 375x1242: PGD iterations per second end to end on the device

  compose   training-batch compositing on the device (next-2): B=32 raw 8-bit 1242x375 stereo pairs -> composited
            adversarial / benign frames -> Pillow-exact Lanczos pyramids -> fp32 (loader.AdvBatchComposer); the
            installed Pillow resizing the same pyramids on one host core is timed beside it

usage: python bench_extra.py [--steps K] [--warmup W] [--only dh,md_f2,costvol,compose]
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
    ap.add_argument("--only", default="cfg1,dh,md_f2,costvol,compose,attack")
    args = ap.parse_args()
    from depthmodelhardening_b200 import _lib, synth
    lib = _lib.load()
    dev = torch.device("cuda:0")
    peak = peak_gbs()
    want = args.only.split(",")

    if "cfg1" in want:
        from depthmodelhardening_b200 import objective
        B, H, W = 4, 192, 640
        pb = synth.photo_batch(batch=B, height=H, width=W, frame_ids=(0, "s"), seed=4).to(dev)
        disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}

        def eager():
            for d in disps.values():
                d.grad = None
            losses, _ = objective.photometric_losses(pb.color, disps, pb.K, pb.inv_K, pb.T, pb.frame_ids, pb.scales, H, W,
                                                     noise=pb.noise)
            losses["loss"].backward()
        ms_e = timed(eager, 10 * args.steps, args.warmup)
        g = objective.GraphedObjective(pb.color, pb.disp, pb.K, pb.inv_K, pb.T, pb.frame_ids, pb.scales, H, W,
                                       noise=pb.noise)
        ms_g = timed(lambda: g.graph.replay(), 10 * args.steps, args.warmup)
        ms_c = timed(lambda: g(pb.color, pb.disp, None, None, None, None), 10 * args.steps, args.warmup)
        print(json.dumps({"workload": "configs[0]: photometric objective fwd+bwd, B=4 640x192 stereo", "B": B, "H": H,
                          "W": W, "eager_ms": ms_e, "graph_replay_ms": ms_g, "graph_with_input_copies_ms": ms_c,
                          "mpix_per_s_eager": B * H * W / ms_e / 1e3, "mpix_per_s_graph": B * H * W / ms_g / 1e3,
                          "note": "eager is launch-bound at this size (~25 launches per step); "
                                  "objective.GraphedObjective replays one captured CUDA graph"}))

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
                          "note": "fused: one dmh_photo_scale_dh launch per scale (single-source fast kernel with the depth-hints decision, packed source)"}))

    if "md_f2" in want:
        from depthmodelhardening_b200 import objective, ops
        # BASELINE config 5 (two temporal sources, pose gradients, resolution sweep) and the "true mono+stereo" variant of
        # config 2 (F = 3, [0,-1,1,'s']), each through the multi-source tile kernel (photo_mf.cu, one
        # launch for all sources and scales) and through the general per-scale kernel it replaces (photo_objective.cu)
        for (B, H, W, fids, bpp, label) in ((16, 320, 1024, (0, -1, 1), 208.0, "config 5"),
                                            (8, 480, 1536, (0, -1, 1), 208.0, "config 5"),
                                            (4, 640, 2048, (0, -1, 1), 208.0, "config 5"),
                                            (32, 320, 1024, (0, -1, 1, "s"), 264.0, "config 2, mono+stereo")):
            pb = synth.photo_batch(batch=B, height=H, width=W, frame_ids=fids, seed=6).to(dev)
            disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
            Ts = {k: v.clone().requires_grad_(k != "s") for k, v in pb.T.items()}

            def step():
                for d in disps.values():
                    d.grad = None
                for t in Ts.values():
                    t.grad = None
                losses, _ = objective.photometric_losses(pb.color, disps, pb.K, pb.inv_K, Ts, pb.frame_ids, pb.scales,
                                                         H, W, noise=pb.noise)
                losses["loss"].backward()
            for mf in (True, False):
                old = ops.MULTISOURCE
                ops.MULTISOURCE = mf
                try:
                    ms = timed(step, args.steps, args.warmup)
                finally:
                    ops.MULTISOURCE = old
                print(json.dumps({"workload": "photometric objective fwd+bwd, %d sources %s + pose gradients (%s)"
                                              % (len(fids) - 1, list(fids[1:]), label), "B": B, "H": H, "W": W,
                                  "ms_per_step": ms, "mpix_per_s": B * H * W / ms / 1e3, "algorithmic_bytes_per_px": bpp,
                                  "hbm_frac": bpp * B * H * W / (ms * 1e-3) / 1e9 / peak,
                                  "note": "photo_mf_kernel: all sources and scales in one launch (multi-source tile kernel)"
                                          if mf else "photo_scale_kernel<F> per scale (general multi-source kernel)"}))
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


    if "compose" in want:
        compose_bench(dev, args, lib)

    if "attack" in want:
        attack_context(dev, args)


class _DepthNetR18(torch.nn.Module):
    """Stock PyTorch stand-in for monodepth2's ResNet-18 encoder + DepthDecoder (random init, outside the
    graft): torchvision resnet18 trunk, 5 up-convolution stages with skip connections, sigmoid disparity."""

    def __init__(self):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet18(weights=None)
        self.stem = torch.nn.Sequential(r.conv1, r.bn1, r.relu)
        self.pool, self.l1, self.l2, self.l3, self.l4 = r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4
        enc, dec = [64, 64, 128, 256, 512], [16, 32, 64, 128, 256]
        conv = lambda i, o: torch.nn.Sequential(torch.nn.ReflectionPad2d(1), torch.nn.Conv2d(i, o, 3), torch.nn.ELU())
        self.up0 = torch.nn.ModuleList([conv(enc[4] if i == 4 else dec[i + 1], dec[i]) for i in range(5)])
        self.up1 = torch.nn.ModuleList([conv(dec[i] + (enc[i - 1] if i > 0 else 0), dec[i]) for i in range(5)])
        self.head = torch.nn.Sequential(torch.nn.ReflectionPad2d(1), torch.nn.Conv2d(dec[0], 1, 3))

    def forward(self, x):
        import torch.nn.functional as F
        f0 = self.stem((x - 0.45) / 0.225)
        f1 = self.l1(self.pool(f0)); f2 = self.l2(f1); f3 = self.l3(f2); f4 = self.l4(f3)
        feats, y = [f0, f1, f2, f3, f4], f4
        for i in range(4, -1, -1):
            y = F.interpolate(self.up0[i](y), scale_factor=2, mode="nearest")
            if i > 0:
                y = torch.cat([y, feats[i - 1]], 1)
            y = self.up1[i](y)
        return torch.sigmoid(self.head(y))


def _write_calib():
    import tempfile
    tmp = tempfile.mkdtemp(prefix="dmh_calib_")
    os.makedirs(os.path.join(tmp, "training", "calib"))
    path = os.path.join(tmp, "training", "calib", "003086.txt")
    with open(path, "w") as f:
        p2 = " ".join(repr(v) for v in __import__("depthmodelhardening_b200.patch_ops", fromlist=["x"]).KITTI_P2_003086)
        for k in ("P0", "P1", "P2", "P3"):
            f.write("%s: %s\n" % (k, p2))
        f.write("R0_rect: 1 0 0 0 1 0 0 0 1\nTr_velo_to_cam: 1 0 0 0 0 1 0 0 0 0 1 0\nTr_imu_to_velo: 1 0 0 0 0 1 0 0 0 0 1 0\n")
    return tmp, path


def compose_bench(dev, args, lib):
    """next-2: MonoDataset.prep_adv_data + preprocess for a collated batch on the device."""
    import time
    import numpy as np
    from depthmodelhardening_b200 import loader, synth
    B, H, W, S = 32, 320, 1024, 4
    _, calib = _write_calib()
    pt = synth.patch_batch(batch=1, seed=3).to(dev)
    comp = loader.AdvBatchComposer(pt.obj, pt.mask, {"path": calib}, H, W, S)
    comp.update_adv_obj(synth.rand(tuple(pt.obj.shape), 78).to(dev))
    c0 = synth.frames_u8(41, batch=B).to(dev)
    cs = synth.frames_u8(42, batch=B).to(dev)
    sides = ["l" if i % 2 == 0 else "r" for i in range(B)]
    flips = [(i // 2) % 2 == 1 for i in range(B)]
    z0 = [5.0 + 0.125 * i for i in range(B)]
    al = [-30.0 + 1.875 * i for i in range(B)]
    fn = lambda: comp(c0, cs, sides, flips, z0, al)
    n0 = lib.dmh_launch_count()
    fn()
    launches = lib.dmh_launch_count() - n0
    ms = timed(fn, args.steps, args.warmup)
    pyr = lambda: [loader.pyramid_u8(x, H, W, S) for x in (c0, cs, c0)]
    ms_pyr = timed(pyr, args.steps, args.warmup)
    # the same three pyramids per item with the installed Pillow on one host core (what a DataLoader worker does)
    from PIL import Image
    host = np.ascontiguousarray(np.transpose(c0[0].cpu().numpy(), (1, 2, 0)))
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps * 3):
        im = Image.fromarray(host)
        for i in range(S):
            im = im.resize((W >> i, H >> i), Image.LANCZOS)
    cpu_ms_item = (time.perf_counter() - t0) / reps * 1e3
    # bytes: 2 raw frames in, 3 composites + mask plane out and in again, 3 pyramids (u8 out, in again as the next
    # level's input, fp32 out), the horizontal-pass intermediates out / in
    px, pyr_px = 375 * 1242, sum((H >> i) * (W >> i) for i in range(S))
    bytes_item = 3 * px * (2 + 3 * 2) + 2 * px + 3 * 3 * pyr_px * (2 + 4) + 2 * 3 * 3 * 375 * W
    print(json.dumps({"workload": "training-batch compositing on the device (next-2: prep_adv_data + preprocess)",
                      "B": B, "native": [375, 1242], "H": H, "W": W, "scales": S, "ms_per_batch": ms,
                      "items_per_s": B / (ms * 1e-3), "pyramids_only_ms": ms_pyr, "gpu_launches_per_batch": int(launches),
                      "approx_bytes_per_batch": bytes_item * B, "approx_gbs": bytes_item * B / (ms * 1e-3) / 1e9,
                      "cpu_pillow_pyramids_ms_per_item": cpu_ms_item, "cpu_cores": 1,
                      "note": "GPU: 2 fused warp+composite launches (the fp32 canvases are never materialised), the three "
                              "composites resized as one stack (2 passes per level, to_tensor fused into the second) + the mask image; CPU figure: only the three 4-level Pillow pyramids of one item on one core "
                              "(the reference additionally warps and composites on the CPU, ~1 s per item)"}))


def attack_context(dev, args):
    import random
    import tempfile
    import time
    import numpy as np
    from depthmodelhardening_b200 import attacks, synth
    torch.manual_seed(0)
    net = _DepthNetR18().to(dev).eval()
    tmp = tempfile.mkdtemp(prefix="dmh_calib_")
    os.makedirs(os.path.join(tmp, "training", "calib"))
    with open(os.path.join(tmp, "training", "calib", "003086.txt"), "w") as f:
        p2 = " ".join(repr(v) for v in __import__("depthmodelhardening_b200.patch_ops", fromlist=["x"]).KITTI_P2_003086)
        for k in ("P0", "P1", "P2", "P3"):
            f.write("%s: %s\n" % (k, p2))
        f.write("R0_rect: 1 0 0 0 1 0 0 0 1\nTr_velo_to_cam: 1 0 0 0 0 1 0 0 0 0 1 0\nTr_imu_to_velo: 1 0 0 0 0 1 0 0 0 0 1 0\n")
    attacks.object_dataset_root = tmp
    Ba = 32
    pt = synth.patch_batch(batch=Ba, seed=3).to(dev)
    # 32 placements per iteration without replacement need >= 32 distances / angles (physicalTrans.py:146-155)
    dist = [5 + 0.2 * i for i in range(40)]
    for name, make in (("l0", lambda: attacks.Phy_obj_atk_l0(net, pt.obj, pt.mask, adam_lr=0.5, steps=5, mask_wt=0.06,
                                                            l0_thresh=0.1, dist_range=dist)),
                       ("linf", lambda: attacks.Phy_obj_atk(net, pt.obj, pt.mask, eps=0.1, alpha=0.02, steps=10,
                                                            random_start=False, dist_range=dist))):
        atk = make()
        for tr in (atk.phy_trans_adv, atk.phy_trans_ben):
            tr.angle_range = [-30 + 1.5 * i for i in range(41)]
        random.seed(0); np.random.seed(0)
        atk(pt.scenes, Ba)                                   # warm-up (cuDNN autotune, homography cache)
        torch.cuda.synchronize()
        random.seed(1); np.random.seed(1)
        t0 = time.perf_counter()
        atk(pt.scenes, Ba)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        iters = 10 if name == "l0" else 10
        print(json.dumps({"workload": "CONTEXT: whole %s attack call incl. a stock ResNet-18 depth net fwd+bwd per "
                                      "iteration (network outside the graft)" % name, "attack_batch": Ba,
                          "iterations": iters, "s_per_call": dt, "pgd_iterations_per_s": iters / dt,
                          "note": "wall clock around one forward() of the drop-in class, device-synchronised; "
                                  "the graft's own share per iteration is bench.py stages.patch_pgd_ms"}))


if __name__ == "__main__":
    main()
