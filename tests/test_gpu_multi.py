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


def _peer_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from depthmodelhardening_b200 import dist as D
        from depthmodelhardening_b200 import patch_ops
        out = {"available": D.PeerReducer.available()}
        n = 3 * 260 * 300 + 1                       # the patch gradient + the scalar attack loss (n % 4 == 1)
        red = D.PeerReducer(n, dev)
        gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
        ok_sum, ok_same = True, True
        for step in range(5):                       # consecutive steps: the device-side step counter / flag protocol
            data = torch.randn(n, generator=gen).to(dev)
            red.buffer.copy_(data)
            got = red.allreduce(average=True).clone()
            gathered = [torch.empty_like(data) for _ in range(world)]
            dist.all_gather(gathered, data)
            ref = gathered[0].clone()
            for r in range(1, world):
                ref = ref + gathered[r]             # rank order, fp32
            ref = ref * (1.0 / world)
            ok_sum = ok_sum and torch.equal(got, ref)
            g2 = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(g2, got)
            ok_same = ok_same and all(torch.equal(g2[0], t) for t in g2)
        out["sum_bit_exact"] = ok_sum
        out["identical_across_ranks"] = ok_same
        # fused L-inf update == dmh_pgd_linf_step on the reduced gradient, bit for bit
        numel = n - 1
        adv = torch.rand(numel, generator=gen).to(dev)
        clean = torch.rand(numel, generator=gen).to(dev)
        dist.broadcast(adv, 0); dist.broadcast(clean, 0)
        data = (torch.randn(n, generator=gen) * (torch.rand(n, generator=gen) > 0.3)).to(dev)   # exact zeros: sign(0)
        red.buffer.copy_(data)
        adv_out = torch.empty_like(adv)
        g = red.allreduce(average=True, linf=(adv, clean, 0.02, 0.1, adv_out)).clone()
        ref_adv = patch_ops.pgd_linf_step(adv.view(1, 3, 260, 300), g[:numel].view(1, 3, 260, 300),
                                          clean.view(1, 3, 260, 300), alpha=0.02, eps=0.1).reshape(-1)
        out["fused_linf_bit_exact"] = torch.equal(adv_out, ref_adv)
        # CUDA-graph capture and replay (three replays = three more steps of the protocol)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            red.buffer.copy_(data)
            red.allreduce(average=False)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            red.buffer.copy_(data)
            res = red.allreduce(average=False)
        ok_graph = True
        for _ in range(3):
            graph.replay()
            torch.cuda.synchronize(dev)
            gathered = [torch.empty_like(data) for _ in range(world)]
            dist.all_gather(gathered, data)
            ref = gathered[0].clone()
            for r in range(1, world):
                ref = ref + gathered[r]
            ok_graph = ok_graph and torch.equal(res, ref)
        out["graph_replay_bit_exact"] = ok_graph
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
def test_peer_memory_allreduce_matches_rank_order_sum():
    """dmh_peer_allreduce (one kernel over NVLink peer memory, no NCCL): bit-exact against the rank-order fp32 sum of
    the ranks' buffers over five consecutive steps, identical on every rank, the fused L-inf update equal to
    dmh_pgd_linf_step on the reduced gradient, and replayable from a CUDA graph."""
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert r in ret, "rank %d did not finish" % r
        for k, v in ret[r].items():
            assert v, (r, k)
