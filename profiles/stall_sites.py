#!/usr/bin/env python
"""Where the warps wait: per-SASS-instruction stall samples from an `ncu --set full --import-source on` report.
Prints the totals per stall reason and the top instructions for the chosen reasons.
usage: python profiles/stall_sites.py X.ncu-rep [reason=long_sb] [top=25] [kernel-id]"""
import csv
import subprocess
import sys


def main(path, reason="long_sb", top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    recs = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        recs.append(r)
    reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: sum(int(r[hdr[n]] or 0) for r in recs) for n in reasons}
    all_s = sum(int(r[hdr["# Samples"]] or 0) for r in recs)
    inst = sum(int(r[hdr["Instructions Executed"]] or 0) for r in recs)
    print("# %s: %d SASS instructions, %d warp-instructions executed, %d samples" % (path, len(recs), inst, all_s))
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print("  %-28s %8d  %5.1f %%" % (n, v, 100.0 * v / max(all_s, 1)))
    col = hdr["stall_" + reason]
    print("# top instructions by stall_%s (index = position in the kernel)" % reason)
    order = sorted(range(len(recs)), key=lambda i: -int(recs[i][col] or 0))[:top]
    for i in order:
        r = recs[i]
        print("  %5d  %7s samples  exec %9s  %s" % (i, r[col], r[hdr["Instructions Executed"]], r[hdr["Source"]].strip()[:100]))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], a[1] if len(a) > 1 else "long_sb", int(a[2]) if len(a) > 2 else 25)
