"""Host-side data-parallel logic on CPU: study sharding == DistributedSampler(shuffle=False), and the reward /
baseline gather with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cxrmate_b200 import sharding


@pytest.mark.parametrize("n,world", [(1, 1), (7, 2), (8, 2), (9, 4), (3, 8), (64, 8), (257, 8)])
def test_shard_matches_distributed_sampler(n, world):
    ds = list(range(n))
    seen = []
    for r in range(world):
        ref = list(torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=r, shuffle=False))
        got = sharding.shard_studies(n, r, world)
        assert got == ref
        seen += got
    assert set(seen) == set(ds)


def test_shard_edge_cases():
    assert sharding.shard_studies(0, 0, 4) == []
    with pytest.raises(ValueError):
        sharding.shard_studies(4, 4, 4)
    assert sharding.batches([0, 2, 4, 6, 8], 2) == [[0, 2], [4, 6], [8]]


def test_gather_without_process_group_is_identity():
    r, b = torch.arange(4.0), torch.ones(4)
    R, B = sharding.gather_rewards(r, b)
    assert R.shape == (1, 4) and torch.equal(R[0], r) and torch.equal(B[0], b)
    with pytest.raises(ValueError):
        sharding.gather_rewards(torch.zeros(3), torch.zeros(4))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_studies(n, rank, world)
        # stand-in for the engine: a deterministic per-study "reward" and "baseline"
        reward = torch.tensor([0.25 * i - 1.0 for i in mine])
        baseline = torch.tensor([0.5 * i + 2.0 for i in mine])
        R, B = sharding.gather_rewards(reward, baseline)
        ms = sharding.max_over_ranks(10.0 + rank, torch.device("cpu"))
        q.put((rank, sharding.unshard(R, n).tolist(), sharding.unshard(B, n).tolist(), ms))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_gather_rewards_world2_gloo(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=60) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_r = [0.25 * i - 1.0 for i in range(n)]
    want_b = [0.5 * i + 2.0 for i in range(n)]
    for rank, r, b, ms in outs:
        assert r == want_r and b == want_b          # every rank sees the whole batch in dataset order
        assert ms == 11.0                           # max over ranks


# ------------------------------------------------------------------------------ generated-prompt subject schedule
def _schedule_cases():
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "subject_schedule.json")
    return json.load(open(path))


def test_allocate_subjects_matches_reference_golden():
    """tests/golden/subject_schedule.json was written by the real PreviousReportSubset.allocate_subjects_to_rank
    (tests/golden/gen_subject_schedule.py): same served order for every world size / mini-batch / shuffle setting"""
    from cxrmate_b200 import sharding as S
    cases = _schedule_cases()
    assert len(cases) >= 20
    for c in cases:
        lists = S.subject_study_lists(c["subject_ids"], c["study_ids"])
        got = S.allocate_subjects(lists, c["world"], c["mbatch"], seed=c["seed"], shuffle_subjects=c["shuffle"])
        assert got == c["examples"], (c["world"], c["mbatch"], c["seed"], c["shuffle"])


def test_subject_stays_on_one_lane_in_order():
    """the property the schedule exists for: all studies of a subject are served to one (rank, slot), in consecutive
    batches and in the subject's own order, so that study k's greedy report can prompt study k+1"""
    from cxrmate_b200 import sharding as S
    for c in _schedule_cases():
        world, mb = c["world"], c["mbatch"]
        lists = S.subject_study_lists(c["subject_ids"], c["study_ids"])
        order = S.allocate_subjects(lists, world, mb, seed=c["seed"], shuffle_subjects=c["shuffle"])
        where = {}
        for p, study in enumerate(order):
            where.setdefault(study, S.lane_of_position(p, world, mb))           # an oversampled study: its first visit
        if c.get("oversampled"):
            assert sorted(set(order)) == sorted(s for l in lists for s in l)     # every study, some repeated
            assert len(order) % (world * mb) == 0
        else:
            assert sorted(order) == sorted(s for l in lists for s in l)          # every study exactly once
        for studies in lists:
            lanes = [where[s] for s in studies]
            assert len({(r, slot) for r, _, slot in lanes}) == 1
            batches_ = [b for _, b, _ in lanes]
            assert batches_ == list(range(batches_[0], batches_[0] + len(studies)))
        # and the rank-local view agrees with shard_studies
        for r in range(world):
            mine = [(i, order[i]) for i in S.shard_studies(len(order), r, world)]
            first_visit = {st: order.index(st) for _, st in mine}
            assert all(where[st][0] == r for i, st in mine if first_visit[st] == i)   # (oversampled copies may land elsewhere)


def test_allocate_subjects_oversamples_like_the_reference():
    """non-divisible study count: the reference appends its last (shortest) subject to ONE lane until the count
    divides (data/prompt.py:183-198); lanes that still differ in length lose studies in zip() and its final assert
    fails - same here"""
    from cxrmate_b200 import sharding as S
    assert any(c.get("oversampled") for c in _schedule_cases())
    # 5 studies on 2 lanes: [1,2,3] | [4], [5] -> pad lane 1 with [5] -> [1,2,3] | [4,5,5]
    assert S.allocate_subjects([[1, 2, 3], [4], [5]], world=2, mbatch=1, shuffle_subjects=False) == [1, 4, 2, 5, 3, 5]
    with pytest.raises(AssertionError):
        S.allocate_subjects([[1, 2, 3], [4]], world=2, mbatch=1, shuffle_subjects=False)
    with pytest.raises(ValueError):
        S.allocate_subjects([[1]], world=0, mbatch=1)


def test_balance_by_images_levels_the_ranks():
    """bench batch of the 8-GPU run (make_images(seed=1234+rank) image counts): DistributedSampler order leaves ranks
    with 80..102 images; the balanced deal keeps 32 studies per rank and levels the image counts to within one study"""
    import torch
    from cxrmate_b200 import sharding as S
    counts = []
    for rank in range(8):
        g = torch.Generator().manual_seed(1234 + rank)
        torch.randn(1, generator=g)
        counts += torch.randint(1, 6, (32,), generator=g).tolist()
    parts = S.balance_by_images(counts, 8)
    assert sorted(i for p in parts for i in p) == list(range(256))
    assert all(len(p) == 32 for p in parts)
    loads = [sum(counts[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 1, loads
    naive = [sum(counts[r * 32:(r + 1) * 32]) for r in range(8)]
    assert max(loads) < max(naive)
    with pytest.raises(ValueError):
        S.balance_by_images([1, 2, 3], 2)


def _allreduce_worker(rank, world, port, q):
    import os

    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cxrmate_b200 import training

        class FakeEngine:
            """stands in for the CUDA engine: stage s writes (rank + 1) * (s + 1) into the slots of stage s"""
            train_stages = 4

            def grad_layout(self, lora_only):
                return [("a", 0, 3, 0), ("b", 3, 2, 1), ("c", 5, 4, 1), ("d", 9, 1, 3)]     # stage 2 owns nothing

            def train_step(self, *tensors, grads=None, stage=-1, _keep=None, **kw):
                if grads is None:
                    grads = torch.zeros(10)
                for _, o, ne, s_ in self.grad_layout(False):
                    if stage in (-1, s_):
                        grads[o:o + ne] = (rank + 1) * (s_ + 1)
                self._train_keep = {"loss": torch.tensor([float(rank)])}
                return self._train_keep["loss"][0], grads

        t = torch.zeros(2, 8, dtype=torch.int64)
        _, g = training._staged(FakeEngine(), (t, t, t, t, t), dict(loss_kind="ce", ignore_index=4, lora_only=False), None, True)
        q.put((rank, g.tolist()))
    finally:
        dist.destroy_process_group()


def test_staged_gradient_all_reduce_gloo_world2():
    """the bucket-by-bucket gradient all-reduce of cxrmate_b200.training (NCCL on the GPUs; gloo here): every stage's
    slice is averaged over the ranks, stages without slots are skipped"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29671
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    mean = 1.5                                                           # mean of (rank + 1) over ranks 0, 1
    want = [mean * 1] * 3 + [mean * 2] * 6 + [mean * 4]
    assert res[0] == pytest.approx(want) and res[1] == pytest.approx(want)
