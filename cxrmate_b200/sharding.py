"""Data-parallel plumbing of the SCST rollout step.

The rollout shards by STUDY with no data-path collective: every rank owns a full
replica of the engine (weights 0.44 GB) and runs encode -> rollouts -> rewards on
its own studies.  The only exchange of the path is the per-step gather of the
`reward` / `baseline` vectors ([2, B] fp32 per rank) that the reference logs
(reference modules/lightning_modules/longitudinal/scst/gen_prompt.py:252-257).

`shard_studies` restates torch.utils.data.DistributedSampler(shuffle=False,
drop_last=False), which is what the reference's train_dataloader uses
(scst/gen_prompt.py:118): the index list is padded by wrapping around so that it
divides evenly, and rank r takes indices r, r + world, r + 2*world, ...
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_studies(n_studies: int, rank: int, world: int) -> List[int]:
    """Study indices of `rank` (DistributedSampler(shuffle=False) order, wrap-around padding)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_studies <= 0:
        return []
    per_rank = math.ceil(n_studies / world)
    total = per_rank * world
    idx = list(range(n_studies))
    pad = total - n_studies
    if pad:
        idx += (idx * math.ceil(pad / len(idx)))[:pad]
    return idx[rank:total:world]


def batches(indices: List[int], batch: int) -> List[List[int]]:
    """DataLoader(batch_size=batch, drop_last=False) over a rank's index list."""
    return [indices[i:i + batch] for i in range(0, len(indices), batch)]


def gather_rewards(reward: torch.Tensor, baseline: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-rank reward / baseline vectors in one collective.

    reward, baseline: [B] fp32 on the rank's device (same B on every rank).
    Returns ([world, B], [world, B]); row r holds rank r's values.  Without an initialised process group
    (single GPU) the inputs are returned with a leading axis of 1 and nothing is communicated."""
    if reward.shape != baseline.shape or reward.dim() != 1:
        raise ValueError("reward and baseline must be 1-D tensors of the same length")
    if not (dist.is_available() and dist.is_initialized()):
        return reward[None], baseline[None]
    world = dist.get_world_size(group)
    mine = torch.stack((reward.float(), baseline.float()))          # [2, B]
    out = torch.empty((world * 2, mine.shape[1]), dtype=mine.dtype, device=mine.device)   # ranks concatenated on dim 0
    dist.all_gather_into_tensor(out, mine.contiguous(), group=group)
    out = out.view(world, 2, -1)
    return out[:, 0], out[:, 1]


def unshard(per_rank: torch.Tensor, n_studies: int) -> torch.Tensor:
    """Inverse of shard_studies for gathered values: per_rank [world, per_rank_count] -> [n_studies] in dataset
    order (wrap-around duplicates dropped)."""
    world, cnt = per_rank.shape
    flat = per_rank.t().reshape(-1)          # position k*world + r  <-  per_rank[r, k]  ==  dataset index k*world + r
    return flat[:n_studies]


def max_over_ranks(ms: float, device) -> float:
    """Device-timed milliseconds -> max over ranks (bench.py's timing rule)."""
    if not (dist.is_available() and dist.is_initialized()):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
