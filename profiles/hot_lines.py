#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: python profiles/hot_lines.py X.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file = ""
    out = []
    tot_inst = tot_samp = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 8 and r[0].isdigit():
            try:
                inst = int(r[7]); samp = int(r[6]) if r[6] != "-" else 0
            except ValueError:
                continue
            out.append((inst, samp, cur_file, int(r[0]), r[1].strip()[:110]))
            tot_inst += inst; tot_samp += samp
    print("# %s: total warp-instructions %d, stall samples %d" % (path, tot_inst, tot_samp))
    print("# %10s %6s %6s %6s  %s" % ("inst", "inst%", "samp", "samp%", "file:line source"))
    for inst, samp, f, ln, src in sorted(out, key=lambda t: -t[1])[:top]:
        print("%12d %6.2f %6d %6.2f  %s:%d  %s" % (inst, 100.0 * inst / max(tot_inst, 1), samp, 100.0 * samp / max(tot_samp, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
