#!/usr/bin/env python
"""Host -> device copy bandwidth of N ranks at once (one process per GPU, torchrun): what bounds the fp32 end-to-end
number of bench.py at N = 8.  Every rank copies its own pinned buffer (512 MiB, the size of a step's fp32 batch) to its
GPU `--reps` times; rank 0 prints the per-rank and the aggregate rate.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/h2d_probe.py"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n = a.mib << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        worst = float(t.item())
        print(json.dumps({"probe": "h2d", "n_gpus": world, "mib_per_copy": a.mib, "ms_per_copy_max_over_ranks": worst,
                          "gbs_per_rank": n / (worst * 1e-3) / 1e9, "gbs_aggregate": world * n / (worst * 1e-3) / 1e9,
                          "cpus": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
