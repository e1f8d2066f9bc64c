"""Scene sharding across the GPUs of one box.

Scenes are independent (no cross-scene op anywhere in forward_test,
transfuser_model_v2.py:578-641), so a batch is cut into contiguous blocks, one per rank;
every rank runs the head on its block with replicated weights and no data-path
collective.  The single exchange step is the gather of the planned trajectories
(96 B per scene), done with one NCCL all-gather (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; the first ``total % world`` ranks get one
    scene more, so ragged batches need no padding scenes."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_scenes(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None
                  ) -> torch.Tensor:
    """All-gather per-scene results (first dim = local scenes) into the full batch order.

    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest
    shard for the collective and trimmed afterwards.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(total, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} scenes, expected {hi - lo}")
    base, rem = divmod(total, world)
    if rem == 0:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    cap = base + 1
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    buf = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = []
    for r in range(world):
        l, h = shard_bounds(total, r, world)
        parts.append(buf[r * cap: r * cap + (h - l)])
    return torch.cat(parts, dim=0)


class ShardedPlanner:
    """Run a planning head on this rank's block of scenes and gather the trajectories.

    ``head`` is any callable with the ``TrajectoryHead.forward`` contract.  ``keys`` are the
    entries of the output dict to gather (default: only ``"trajectory"``, 96 B per scene).
    """

    def __init__(self, head, group: Optional[dist.ProcessGroup] = None,
                 keys=("trajectory",)):
        self.head = head
        self.group = group
        self.keys = tuple(keys)

    def plan_local(self, ego, agents, bev, noise=None, **kw) -> Dict[str, torch.Tensor]:
        return self.head(ego, agents, bev, noise=noise, **kw)

    def plan(self, ego, agents, bev, total: int, noise=None, **kw) -> Dict[str, torch.Tensor]:
        """Inputs are this rank's shard (see ``shard_bounds``); returns full-batch tensors."""
        out = self.plan_local(ego, agents, bev, noise=noise, **kw)
        return {k: gather_scenes(out[k], total, self.group) for k in self.keys}
