"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

Recipe that makes the UNMODIFIED reference available on the GPU box: copies the Python sources of the hot path
(the files SURVEY.md section 8 cites and what they import) from /root/reference, where they lie, into
`oracle/_ref/reference/` -- git-ignored (the history stays free of reference code) but not gpurun-ignored, so it
travels with the snapshot like the built `.so`.  `__graft_entry__.build()` runs it when /root/reference is present.

What the copy is used for (oracle/ref_step.py): `bench.py --impl reference` (the reference's own code on the host
cores, `cpu_baseline.kind == "reference"`), `bench.py`'s `gpu_eager_baseline` (the same code on `cuda`: the stock
eager-PyTorch path on the same B200) and the drop-in integration tests on hardware.
    python -m oracle.build_ref
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DMH_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref", "reference")

# python sources only: no assets, splits, notebooks or images
SUBTREES = ("", "torchattacks", "torchattacks/attacks", "preprocessing", "DepthNetworks/monodepth2",
            "DepthNetworks/monodepth2/networks", "DepthNetworks/monodepth2/datasets",
            "DepthNetworks/depth-hints", "DepthNetworks/depth-hints/networks", "DepthNetworks/depth-hints/datasets",
            "DepthNetworks/manydepth2/manydepth", "DepthNetworks/manydepth2/manydepth/networks",
            "DepthNetworks/manydepth2/manydepth/datasets")


def build(verbose: bool = True) -> int:
    if not os.path.isdir(SRC):
        if verbose:
            print("build_ref: %s not present (GPU box): using the prebuilt copy" % SRC)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for sub in SUBTREES:
        s = os.path.join(SRC, sub)
        if not os.path.isdir(s):
            continue
        d = os.path.join(DST, sub)
        os.makedirs(d, exist_ok=True)
        for name in sorted(os.listdir(s)):
            if name.endswith(".py"):
                shutil.copyfile(os.path.join(s, name), os.path.join(d, name))
                n += 1
    if verbose:
        print("build_ref: %d reference sources -> %s" % (n, os.path.relpath(DST, ROOT)))
    return n


if __name__ == "__main__":
    sys.exit(0 if build() >= 0 else 1)
