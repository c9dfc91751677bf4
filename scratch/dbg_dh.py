import sys, torch, numpy as np
sys.path.insert(0, ".")
from depthmodelhardening_b200 import synth
from oracle import depth_hints as OD, photometric as OP
from oracle.make_golden_dh import CASES
from tests.test_depth_hints import _run_cuda, _opts
dev = torch.device("cuda:0")
for name in ("avg_hints", "stereo_hints"):
    skw, use_hints, over = CASES[name]
    pb = synth.photo_batch(**skw)
    total, ref_losses, grads, aux0 = OD.objective_from_batch(pb, use_hints, _opts(pb, over), return_aux=True)
    losses, aux, disps = _run_cuda(pb, dev, use_hints, over)
    for s in pb.scales:
        sel = aux[("argmin", s)].cpu().numpy(); ref = aux0[("argmin", s)][:, 0].numpy()
        print(name, s, "flips", int((sel != ref).sum()), "of", sel.size,
              "reproj", float(losses["reproj_loss/%d" % s]), float(ref_losses["reproj_loss/%d" % s]),
              "hint", float(losses["depth_hint_loss/%d" % s]), float(ref_losses["depth_hint_loss/%d" % s]),
              "n_hint", int((ref == 2).sum()))
