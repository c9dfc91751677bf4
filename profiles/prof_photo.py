#!/usr/bin/env python
"""Stage 2 only (the photometric objective fwd+bwd at the bench configuration), a few steps: the command ncu wraps to
capture the photometric kernels.  usage: python profiles/prof_photo.py [--multiscale 0|1] [--steps 2] [--batch 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--multiscale", type=int, default=1)
    a = ap.parse_args()
    from depthmodelhardening_b200 import ops
    ops.MULTISCALE = bool(a.multiscale)
    pb, _ = bench.make_host_workload(a.batch, 0, False)
    s2 = bench.Stage2(pb, torch.device("cuda:0"))
    for _ in range(a.steps):
        s2.step()
    torch.cuda.synchronize()
    print("ms/step %.3f" % bench.timed_loop(s2.step, 10, 2, 1))


if __name__ == "__main__":
    main()
