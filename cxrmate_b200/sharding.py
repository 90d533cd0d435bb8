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
from typing import List, Sequence, Tuple

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


def balance_by_images(n_images: Sequence[int], world: int) -> List[List[int]]:
    """Study indices per rank for one global SCST batch, balanced by WORK rather than by count.

    Encoder and cross-attention cost grow with the number of valid images of a study (1..5 in MIMIC-CXR), so
    `DistributedSampler` order gives ranks 80..102 images for the same 32 studies and the step time is the slowest
    rank's.  Every rank still gets exactly len(n_images) / world studies (the rollout batch shape is fixed); studies
    are dealt largest-first to the rank with the fewest images among those that still have room (LPT with a
    cardinality constraint).  Deterministic; ties go to the lower rank, equal studies keep dataset order.
    Independent studies only: the generated-prompt schedule keeps subjects on their lane (allocate_subjects)."""
    n = len(n_images)
    if world < 1 or n % world != 0:
        raise ValueError("the global batch must divide evenly over the ranks")
    per = n // world
    order = sorted(range(n), key=lambda i: (-int(n_images[i]), i))
    load, out = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(out[r]) < per), key=lambda r: (load[r], r))
        out[r].append(i)
        load[r] += int(n_images[i])
    return [sorted(x) for x in out]


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


# ---------------------------------------------------------------------------------------------------------------
# Generated-prompt SCST: subjects, not studies, are the unit of sharding.
# Study k's greedy report is study k+1's prompt, so all studies of a subject must visit ONE (rank, batch slot) in
# chronological order.  Restates PreviousReportSubset.allocate_subjects_to_rank (reference data/prompt.py:142-213),
# pinned against the real method by tests/golden/subject_schedule.json.
# ---------------------------------------------------------------------------------------------------------------
def subject_study_lists(subject_ids, study_ids) -> List[List[int]]:
    """`df.drop_duplicates(subset=['study_id']).groupby('subject_id')['study_id'].apply(list).tolist()`
    (data/prompt.py:160-162) without pandas: subjects in ascending id order, each subject's studies in first-occurrence
    order of the table, every study once."""
    seen, per_subject = set(), {}
    for subj, study in zip(subject_ids, study_ids):
        if study in seen:
            continue
        seen.add(study)
        per_subject.setdefault(subj, []).append(study)
    return [per_subject[s] for s in sorted(per_subject)]


def allocate_subjects(subject_lists: List[List[int]], world: int, mbatch: int, seed=None,
                      shuffle_subjects: bool = True) -> List[int]:
    """Order in which the dataset serves study ids so that, with DistributedSampler(shuffle=False) and
    batch_size=mbatch, every subject stays on one (rank, batch slot) lane with its studies in consecutive batches.

    Same algorithm as the reference (data/prompt.py:164-213): subjects sorted by number of studies (stable, largest
    first) are packed greedily onto the currently shortest of world*mbatch lanes (numpy argmin: first minimum); when
    the study count does not divide by world*mbatch the LAST (shortest) subject is appended again - oversampled - to
    the lane that was shortest after packing until it does (the reference tests a length list it never updates, so
    every copy lands on that one lane: reproduced); lanes are optionally shuffled per lane with
    `random.seed(seed); random.sample(...)`, flattened, and interleaved element-wise with zip(), which truncates to
    the shortest lane.  Like the reference, the result must still serve every study: AssertionError otherwise."""
    import itertools
    import random

    if world < 1 or mbatch < 1:
        raise ValueError("world and mbatch must be positive")
    lanes_n = world * mbatch
    lists = sorted(subject_lists, key=len, reverse=True)          # list.sort is stable, like the reference's
    lanes: List[List[List[int]]] = [[] for _ in range(lanes_n)]
    total = [0] * lanes_n
    for studies in lists:
        i = total.index(min(total))                                # np.argmin: first occurrence of the minimum
        lanes[i].append(studies)
        total[i] += len(studies)
    n_served = sum(total)
    if n_served % lanes_n != 0:
        if not lists or not lists[-1]:
            raise ValueError("nothing to oversample")
        i = total.index(min(total))                                # stale in the reference's loop: one fixed lane
        while n_served % lanes_n != 0:
            lanes[i].append(lists[-1])
            n_served += len(lists[-1])
    if shuffle_subjects:
        random.seed(seed)
        flat = [list(itertools.chain(*random.sample(lane, k=len(lane)))) for lane in lanes]
    else:
        flat = [list(itertools.chain(*lane)) for lane in lanes]
    order = [study for row in zip(*flat) for study in row]
    every = {st for studies in subject_lists for st in studies}
    assert set(order) == every, ("interleaving dropped studies: the lanes are not equally long "
                                 f"({min(map(len, flat))}..{max(map(len, flat))}); the reference's final assert fails too")
    return order


def lane_of_position(pos: int, world: int, mbatch: int) -> Tuple[int, int, int]:
    """(rank, batch index on that rank, slot inside the batch) of position `pos` of the served order under
    DistributedSampler(shuffle=False) + DataLoader(batch_size=mbatch)."""
    rank, k = pos % world, pos // world
    return rank, k // mbatch, k % mbatch
