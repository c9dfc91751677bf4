#!/usr/bin/env python
"""Per-kernel average durations of the bench step measured in place (torch.profiler / CUPTI activity records:
warm caches, real launch order), unlike the serialised cold-cache ncu launch list.
usage: python profiles/kernel_times.py [--steps 10] [--batch 32]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--multiscale", type=int, default=1, help="0: one photometric launch per scale (round-1 form)")
    a = ap.parse_args()
    from depthmodelhardening_b200 import ops
    ops.MULTISCALE = bool(a.multiscale)
    dev = torch.device("cuda:0")
    pb, pt = bench.make_host_workload(a.batch, 0, True)
    s2, s1 = bench.Stage2(pb, dev), bench.Stage1(pt, dev, 1)

    def step():
        s1.step()
        s2.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for name, fn in (("stage 1 (patch step)", s1.step), ("stage 2 (objective fwd+bwd)", s2.step), ("step", step)):
        print("%9.1f us  %s, CUDA events over 20 iterations" % (1e3 * bench.timed_loop(fn, 20, 3, 1), name))
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    tot = collections.Counter()
    cnt = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            tot[e.name] += e.device_time
            cnt[e.name] += 1
    s = 0.0
    for k, v in tot.most_common():
        print("%9.1f us/step  %3d launches/step  avg %8.1f us  %s" % (v / a.steps, cnt[k] // a.steps, v / cnt[k], k[:90]))
        s += v / a.steps
    print("%9.1f us/step  sum of kernel time" % s)


if __name__ == "__main__":
    main()
