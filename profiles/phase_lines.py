#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report
(`--page source --print-source cuda,sass`), optionally grouped into line-range phases.
usage: python profiles/phase_lines.py X.ncu-rep KERNEL [file:lo-hi=name ...]"""
import collections
import csv
import subprocess
import sys


def main(path, kernel, groups, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", kernel, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr = "", None
    agg, smp, src = collections.Counter(), collections.Counter(), {}
    for r in csv.reader(raw.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            try:
                inst, s = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            k = (cur, int(r[0]))
            agg[k] += inst; smp[k] += s; src[k] = r[1].strip()[:100]
    tot, ts = sum(agg.values()), sum(smp.values())
    print("# %s %s: warp-instructions %d, stall samples %d" % (path, kernel, tot, ts))
    if groups:
        ph, ps = collections.Counter(), collections.Counter()
        for (f, l), v in agg.items():
            name = f
            for g in groups:
                spec, nm = g.split("=")
                gf, rng = spec.split(":")
                lo, hi = map(int, rng.split("-"))
                if gf == f and lo <= l <= hi:
                    name = nm
            ph[name] += v; ps[name] += smp[(f, l)]
        for k, v in ph.most_common():
            print("%-28s %6.2f%% inst  %6.2f%% samples" % (k, 100 * v / tot, 100 * ps[k] / ts))
    for k, v in agg.most_common(top):
        print("%6.2f%% %6.2f%% %s:%d %s" % (100 * v / tot, 100 * smp[k] / ts, k[0], k[1], src[k]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
