"""Two GPUs, NCCL: the attack classes optimising ONE universal patch over both ranks (opt-in
`enable_patch_sync`).  Needs >= 2 CUDA devices: skipped on a single-GPU box (the 1-GPU suite covers
the kernels; tests/test_dist_gloo.py covers the collectives' logic on CPU).

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -q -m gpu
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, calib_root, ret):
    import random

    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from depthmodelhardening_b200 import attacks, synth
        from depthmodelhardening_b200 import dist as D
        from tests.test_gpu_patch import _tiny
        attacks.object_dataset_root = calib_root
        model = _tiny(dev).eval()
        # the global batch of 4 scenes sharded over the ranks; every rank seeds its RNGs DIFFERENTLY on purpose
        full = synth.patch_batch(batch=4, seed=9)
        lo, hi = D.shard_range(4)
        scenes = full.scenes[lo:hi].to(dev)
        obj, mask = full.obj.to(dev), full.mask.to(dev)
        random.seed(100 + rank)
        np.random.seed(200 + rank)
        torch.manual_seed(300 + rank)
        out = {}
        for name, make in (("linf", lambda: attacks.Phy_obj_atk(model, obj.clone(), mask.clone(), eps=0.1, alpha=0.02,
                                                                steps=3, random_start=True, dist_range=list(range(5, 10, 2)))),
                           ("l0", lambda: attacks.Phy_obj_atk_l0(model, obj.clone(), mask.clone(), adam_lr=0.5, steps=2,
                                                                 mask_wt=0.06, l0_thresh=0.1, dist_range=list(range(5, 10, 2))))):
            atk = make().enable_patch_sync()
            patch = atk(scenes.clone(), hi - lo)[3].contiguous()
            gathered = [torch.empty_like(patch) for _ in range(world)]
            dist.all_gather(gathered, patch)
            out[name + "_identical"] = all(torch.equal(gathered[0], t) for t in gathered)
            out[name + "_moved"] = bool((patch - obj).abs().max() > 0)
            # without the opt-in the ranks (different RNG streams, different scenes) end on different patches and
            # issue no collective at all
            solo = make()
            p2 = solo(scenes.clone(), hi - lo)[3].contiguous()
            g2 = [torch.empty_like(p2) for _ in range(world)]
            dist.all_gather(g2, p2)
            out[name + "_solo_differs"] = not torch.equal(g2[0], g2[1])
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
def test_universal_patch_stays_identical_across_two_gpus():
    import tempfile

    import torch.multiprocessing as mp
    from oracle.refload import write_calib
    root = tempfile.mkdtemp(prefix="dmh_calib_")
    write_calib(root)
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, root, ret), nprocs=2, join=True)
    for r in (0, 1):
        assert r in ret, "rank %d did not finish" % r
        for k, v in ret[r].items():
            assert v, (r, k)
