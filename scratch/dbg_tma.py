import sys, torch
sys.path.insert(0, ".")
from depthmodelhardening_b200 import objective, synth
dev = torch.device("cuda:0")
for (B, H, W) in [(2, 64, 96), (2, 96, 160), (1, 320, 1024), (8, 320, 1024)]:
    pb = synth.photo_batch(batch=B, height=H, width=W, frame_ids=(0, "s"), seed=41).to(dev)
    disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
    losses, _ = objective.photometric_losses(pb.color, disps, pb.K, pb.inv_K, pb.T, pb.frame_ids, pb.scales,
                                             pb.height, pb.width, noise=pb.noise)
    losses["loss"].backward()
    torch.cuda.synchronize()
    print("ok", B, H, W, float(losses["loss"]), flush=True)
