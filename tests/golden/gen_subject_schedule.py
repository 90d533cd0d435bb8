"""Golden vectors for cxrmate_b200.sharding.allocate_subjects, written by the REAL reference method
`PreviousReportSubset.allocate_subjects_to_rank` (/root/reference/data/prompt.py:142-213) called on a stub `self`
(the dataset class itself needs MIMIC-CXR tables that are not available offline).

    python tests/golden/gen_subject_schedule.py        # authoring container only; writes subject_schedule.json
"""
import json
import os
import random
import sys
import types
import warnings

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, "/root/reference")
from data.prompt import PreviousReportSubset  # noqa: E402


def reference_schedule(subject_ids, study_ids, world, mbatch, seed, shuffle):
    stub = types.SimpleNamespace()
    stub.df = pd.DataFrame({"subject_id": subject_ids, "study_id": study_ids})
    stub.use_generated, stub.scst_generated, stub.mbatch_size = True, True, mbatch
    real = torch.distributed.get_world_size
    torch.distributed.get_world_size = lambda *a, **k: world
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            PreviousReportSubset.allocate_subjects_to_rank(stub, seed=seed, shuffle_subjects=shuffle)
    finally:
        torch.distributed.get_world_size = real
    return [int(x) for x in stub.examples]


def make_table(n_subjects, rng, divisible_by=None):
    subject_ids, study_ids, sid = [], [], 50000000
    for s in rng.sample(range(10000000, 10000000 + 10 * n_subjects), n_subjects):
        for _ in range(rng.choice([1, 1, 1, 2, 2, 3, 5, 9])):
            sid += rng.randint(1, 50)
            n_rows = rng.choice([1, 1, 2, 3])            # several images (rows) per study
            subject_ids += [s] * n_rows
            study_ids += [sid] * n_rows
    if divisible_by:
        # The reference interleaves the world * mbatch lists with zip(), i.e. it REQUIRES them to end up equally long
        # (its final assert fails otherwise): as in MIMIC-CXR, plenty of single-study subjects let the greedy packing
        # level out, and the study count is made divisible (the reference's own padding loop for the non-divisible
        # case appends to a fixed list without updating the lengths it tests).
        s = 99000000
        for _ in range(10 * divisible_by):
            sid += 1; s += 1
            subject_ids.append(s); study_ids.append(sid)
        n = len(set(study_ids))
        while n % divisible_by:
            sid += 1; s += 1; n += 1
            subject_ids.append(s); study_ids.append(sid)
    order = list(range(len(study_ids)))
    rng.shuffle(order)                                     # table rows in arbitrary order, as in the real CSVs
    return [subject_ids[i] for i in order], [study_ids[i] for i in order]


def main():
    rng = random.Random(20260117)
    cases = []
    for world, mbatch in [(1, 1), (1, 4), (2, 2), (2, 3), (8, 4)]:
        for n_subjects in (world * mbatch + 3, 3 * world * mbatch + 1):
            for shuffle, seed in [(False, None), (True, 41)]:
                subj, stud = make_table(n_subjects, rng, divisible_by=world * mbatch)
                ex = reference_schedule(subj, stud, world, mbatch, seed, shuffle)
                cases.append(dict(world=world, mbatch=mbatch, seed=seed, shuffle=shuffle, subject_ids=subj, study_ids=stud,
                                  examples=ex))
    # non-divisible study counts: the reference oversamples its last (single-study) subject into one lane.  Only
    # tables the reference itself accepts are kept (its zip() truncates unequal lanes and its final assert then fails).
    n_odd = 0
    for world, mbatch in [(1, 4), (2, 2), (2, 3), (8, 4)]:
        for attempt in range(200):
            subj, stud = make_table(3 * world * mbatch + 1, rng, divisible_by=world * mbatch)
            extra = rng.randint(1, world * mbatch - 1)
            s0, sid0 = max(subj) + 1, max(stud) + 1
            for k in range(extra):
                subj.append(s0 + k); stud.append(sid0 + k)
            assert len(set(stud)) % (world * mbatch) != 0
            try:
                ex = reference_schedule(subj, stud, world, mbatch, 7, True)
            except AssertionError:
                continue
            cases.append(dict(world=world, mbatch=mbatch, seed=7, shuffle=True, subject_ids=subj, study_ids=stud,
                              examples=ex, oversampled=True))
            n_odd += 1
            break
    print("non-divisible cases accepted by the reference:", n_odd)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "subject_schedule.json")
    json.dump(cases, open(out, "w"))
    print(len(cases), "cases ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
