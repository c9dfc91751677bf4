#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel, split at BAR.SYNC (the kernels' phases are fully unrolled, so the
static count per thread is the dynamic count up to the untaken slow paths).
usage: python profiles/sass_phases.py OBJ_OR_SO KERNEL_SUBSTRING [top]"""
import collections
import re
import subprocess
import sys


def main(obj, pat, top=14):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, phase = None, 0
    cnt = collections.defaultdict(lambda: collections.defaultdict(collections.Counter))
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur, phase = m.group(1), 0
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if not m or cur is None or pat not in cur:
            continue
        base = m.group(2).split(".")[0]
        cnt[cur][phase][base] += 1
        if base == "BAR":
            phase += 1
    for fn, ph in cnt.items():
        print("==", fn, "total", sum(sum(c.values()) for c in ph.values()))
        for k, c in ph.items():
            print("  phase %d: %5d  %s" % (k, sum(c.values()), " ".join("%s:%d" % kv for kv in c.most_common(int(top)))))


if __name__ == "__main__":
    main(*sys.argv[1:])
