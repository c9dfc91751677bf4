"""CPU, world_size 2, gloo: the host-side logic of the N>1 path -- batch
sharding, the single fused all-reduce of (patch gradient + scalar loss), and
that the global loss assembled from per-rank numerators equals the
single-process value."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from depthmodelhardening_b200 import dist as D
        from depthmodelhardening_b200 import synth
        from oracle import photometric as OP
        B = 4
        lo, hi = D.shard_range(B)
        assert (hi - lo) == 2 and lo == 2 * rank
        # --- collective 1: shared patch gradient + scalar loss in ONE all-reduce
        g = torch.full((1, 3, 5, 7), float(rank + 1))
        loss = torch.tensor(0.5 * (rank + 1))
        gsum, (lsum,) = D.allreduce_patch_grad(g, [loss], average=False)
        assert torch.equal(gsum, torch.full((1, 3, 5, 7), 3.0)) and float(lsum) == 1.5
        gavg, _ = D.allreduce_patch_grad(g, [], average=True)
        assert torch.equal(gavg, torch.full((1, 3, 5, 7), 1.5))
        # --- collective 2: per-scale loss numerators; denominators static (global B*H*W)
        pb = synth.photo_batch(batch=B, height=32, width=64, frame_ids=(0, "s"), seed=3)
        sh = D.shard_batch({"t": pb.color[(0, 0)], "s": pb.color[("s", 0)], "d": pb.disp[0], "n": pb.noise[0],
                            "K": pb.K, "iK": pb.inv_K, "T": pb.T["s"]}, B)

        def numerator(t, s, d, n, K, iK, T):
            pred, _, _ = OP.warp_from_disp(d, s, K, iK, T, 0.1, 100.0)
            comb = torch.cat((OP.reprojection_loss(s, t) + n, OP.reprojection_loss(pred, t)), 1)
            return torch.min(comb, dim=1)[0].sum()

        local = numerator(sh["t"], sh["s"], sh["d"], sh["n"], sh["K"], sh["iK"], sh["T"]).reshape(1)
        total = D.allreduce_loss_sums(local.clone()) / float(B * 32 * 64)
        full = numerator(pb.color[(0, 0)], pb.color[("s", 0)], pb.disp[0], pb.noise[0], pb.K, pb.inv_K,
                         pb.T["s"]) / float(B * 32 * 64)
        assert abs(float(total) - float(full)) <= 1e-6 * abs(float(full))
        with pytest.raises(ValueError):
            D.shard_range(5)
        # the peer-memory form of collective 1 needs NCCL ranks on NVLink-connected GPUs: under gloo the callers keep
        # the all-reduce above (attacks._sync_patch_grad, bench.Stage1 test `PeerReducer.available()` first)
        assert not D.PeerReducer.available()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok"


def test_single_process_is_identity():
    from depthmodelhardening_b200 import dist as D
    g = torch.ones(2, 3)
    out, sc = D.allreduce_patch_grad(g, [torch.tensor(2.0)])
    assert out is g and float(sc[0]) == 2.0
    assert D.shard_range(8) == (0, 8)


def _sync_worker(rank, world, port, ret):
    """The attack classes' shared-patch collectives are OPT-IN (attacks._sync_patch_grad / _sync_initial_state):
    exercised on CPU tensors through the same helpers the attack loops call (the loops themselves need the kernels)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace
        from depthmodelhardening_b200 import attacks
        gen = torch.Generator().manual_seed(100 + rank)               # every rank draws a DIFFERENT start and gradient
        # --- no sync group (default): purely local, no collective is issued although torch.distributed is up
        local = SimpleNamespace(sync_group=None)
        start = torch.rand(1, 3, 6, 8, generator=gen)
        keep = start.clone()
        attacks._sync_initial_state(local, start)
        g = torch.rand(1, 3, 6, 8, generator=gen)
        assert attacks._sync_patch_grad(local, g) is g and torch.equal(start, keep)
        # --- opted in: rank 0's start everywhere, mean gradient everywhere -> identical sign step on every rank
        synced = SimpleNamespace(sync_group=True)
        pos, neg = torch.rand(1, 3, 6, 8, generator=gen), torch.rand(1, 3, 6, 8, generator=gen)
        attacks._sync_initial_state(synced, pos, neg)
        patch = pos.clone()
        for _ in range(3):                                            # a lock-step "attack": 3 sign steps
            grad = torch.rand(1, 3, 6, 8, generator=gen) - 0.5
            grad = attacks._sync_patch_grad(synced, grad)
            patch = (patch + 0.02 * torch.sign(grad)).clamp(0, 1)
        gathered = [torch.zeros_like(patch) for _ in range(world)]
        dist.all_gather(gathered, patch)
        assert all(torch.equal(gathered[0], t) for t in gathered), "patches diverged across ranks"
        gathered = [torch.zeros_like(neg) for _ in range(world)]
        dist.all_gather(gathered, neg)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_patch_sync_is_opt_in_and_keeps_ranks_identical():
    """without `enable_patch_sync` an attack issues no collective (ordinary DDP training calls the attack
    independently per rank); with it the start is rank 0's and every rank applies the same mean gradient, so the
    shared patch stays bit-identical and both ranks exit cleanly."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sync_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok"
