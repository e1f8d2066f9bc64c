"""CPU, world_size 2 over gloo: the scene-sharding host logic (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusiondrive_b200.parallel import ShardedPlanner, gather_scenes, shard_bounds


def test_shard_bounds_cover_batch():
    for total in (1, 2, 7, 8, 255, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _fake_head(ego, agents, bev, noise=None):
    """Stands in for the CUDA head: a deterministic per-scene function of the inputs."""
    s = ego.sum(dim=(1, 2)) + agents.sum(dim=(1, 2)) + bev.sum(dim=(1, 2, 3))
    traj = s[:, None, None] * torch.arange(24, dtype=torch.float32).view(1, 8, 3)
    return {"trajectory": traj, "mode_idx": (s.abs() * 10).long() % 20}


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        ego = torch.randn(total, 1, 4, generator=g)
        agents = torch.randn(total, 3, 4, generator=g)
        bev = torch.randn(total, 2, 2, 2, generator=g)
        full = _fake_head(ego, agents, bev)
        lo, hi = shard_bounds(total, rank, world)
        planner = ShardedPlanner(_fake_head, keys=("trajectory", "mode_idx"))
        out = planner.plan(ego[lo:hi], agents[lo:hi], bev[lo:hi], total)
        ok = torch.equal(out["trajectory"], full["trajectory"]) and torch.equal(
            out["mode_idx"], full["mode_idx"])
        try:
            gather_scenes(full["trajectory"][: hi - lo + 1], total)
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_plan_world2(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
