#!/usr/bin/env python
"""Dynamic (executed) SASS opcode mix of one kernel from an ncu report, split at BAR.SYNC, in lane-instructions
per pixel.  usage: python profiles/sass_dynamic.py X.ncu-rep KERNEL PIXELS [top]"""
import collections
import csv
import subprocess
import sys


def main(path, kernel, pixels, top=22):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name",
                          kernel, "--launch-count", "1"], capture_output=True, text=True).stdout
    px = float(pixels)
    hdr, phase = None, 0
    inst, smp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == "Address":
            hdr = r
            iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= iI or not r[0].startswith("0x"):
            continue
        parts = r[iS].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        base = op.split(".")[0]
        n = int(r[iI])
        inst[phase] += n; smp[phase] += int(r[iN]); ops[phase][base] += n
        if base == "BAR":
            phase += 1
    tot, ts = sum(inst.values()), max(1, sum(smp.values()))
    for p in sorted(inst):
        print("phase %d: %5.1f%% inst %5.1f%% samples  %6.1f lane-instr/px | %s" % (
            p, 100 * inst[p] / tot, 100 * smp[p] / ts, inst[p] * 32 / px,
            " ".join("%s:%.1f" % (k, v * 32 / px) for k, v in ops[p].most_common(int(top)))))
    print("total lane-instr/px %.1f (warp-instructions %d)" % (tot * 32 / px, tot))


if __name__ == "__main__":
    main(*sys.argv[1:])
