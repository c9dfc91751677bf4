#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary kept under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
       python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep --json KERNEL_SUBSTRING BATCH > profiles/rNN_ncu_<k>.json
         (per-launch DRAM bytes / warp instructions of the first matching launch: what bench.py's roofline.traffic
          quotes, with its source)"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# source: %s (ncu --set full --clock-control none); one block per profiled launch" % path)
    for r in rows[2:]:
        print("\n== %s  [id %s]" % (r[idx["Kernel Name"]][:100], r[idx["ID"]]))
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print("  %-82s %16s %s" % (w, r[idx[w]][:16], units[idx[w]]))


def as_json(path, pat, batch):
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pat in name:
            f = lambda k: float(r[idx[k]].replace(",", ""))
            units = rows[1]
            def byt(k):
                u = units[idx[k]].lower()
                m = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
                return f(k) * m
            out = {"kernel": pat, "kernel_name": name[:160], "batch": int(batch),
                   "dram_bytes_per_launch": byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"),
                   "dram_bytes_read": byt("dram__bytes_read.sum"), "dram_bytes_write": byt("dram__bytes_write.sum"),
                   "warp_instructions_per_launch": f("smsp__inst_executed.sum"),
                   "gpu_time_us_under_ncu": f("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(
                       units[idx["gpu__time_duration.sum"]].lower().replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1.0),
                   "registers_per_thread": f("launch__registers_per_thread"),
                   "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "source": "ncu --set full --clock-control none, %s" % path.split("/")[-1]}
            print(json.dumps(out, indent=1))
            return
    raise SystemExit("no launch matching %r in %s" % (pat, path))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--json":
        as_json(sys.argv[1], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1])
