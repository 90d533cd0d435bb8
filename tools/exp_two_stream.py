"""Experiment: does running the rollout as TWO independent half-batches on two streams (two engines, 16 studies each)
hide the dependency latency of the 71-launch decode chain behind the other half's bandwidth-bound attention?

    python tools/exp_two_stream.py            # prints ms per rollout: one engine B=32 | two engines serial | concurrent
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from cxrmate_b200 import synthetic as S  # noqa: E402
from cxrmate_b200 import synthetic_weights as W  # noqa: E402
from cxrmate_b200.engine import Engine  # noqa: E402


class A:
    studies, images, prompt, tokens = 32, 5, 256, 255


def main():
    a = A()
    dev = torch.device("cuda", 0)
    B, N, T = a.studies, a.images, a.tokens
    nsplit = int(os.environ.get("NSPLIT", "2"))
    counts = bench.global_image_counts(B, N)
    studies = [bench.make_study(a, g, counts[g]) for g in range(B)]
    # deal studies to the halves by image count (serpentine)
    order = sorted(range(B), key=lambda i: -counts[i])
    halves = [[] for _ in range(nsplit)]
    for j, i in enumerate(order):
        k = j % (2 * nsplit)
        halves[k if k < nsplit else 2 * nsplit - 1 - k].append(i)
    print("images per part", [sum(counts[i] for i in h) for h in halves])

    def pack(idx):
        px = torch.stack([studies[i][0] for i in idx]).to(dev)
        P_ = max(len(studies[i][1]) for i in idx)
        pr = torch.full((len(idx), P_), S.PAD, dtype=torch.int64)
        for b, i in enumerate(idx):
            pr[b, : len(studies[i][1])] = studies[i][1]
        return px, pr.to(dev)

    sd, rsd = W.make_cxrmate_weights(seed=0), W.make_cxrbert_weights(seed=1)

    def mk(nb):
        e = Engine(dtype="bf16", max_studies=nb, max_images=N, max_prompt=a.prompt, max_new_tokens=T, rwd_layers=0,
                   enc_chunk=32)
        e.load_state_dict(sd)
        e.finalize()
        return e

    kw = dict(mode="both", max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=S.PAD,
              special_sample=S.SPECIAL_SAMPLE, sections_sample=S.SECTIONS[:3], special_greedy=S.SPECIAL_GREEDY,
              sections_greedy=S.SECTIONS, top_k=50, temperature=1.0)

    full = mk(B)
    px, pr = pack(list(range(B)))
    full.encode(px)
    full.prefill_cross_kv()
    for i in range(2):
        full.rollout(pr, seed=i, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        full.rollout(pr, seed=10 + i, **kw)
    torch.cuda.synchronize()
    print(f"one engine B={B}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per rollout", flush=True)
    del full, px
    torch.cuda.empty_cache()

    engs, ins, streams = [], [], []
    for h in halves:
        e = mk(len(h))
        px_h, pr_h = pack(h)
        e.encode(px_h)
        e.prefill_cross_kv()
        engs.append(e)
        ins.append(pr_h)
        streams.append(torch.cuda.Stream(device=dev))
    torch.cuda.synchronize()
    for e, p in zip(engs, ins):      # serial warm-up: graph capture happens here, one engine at a time
        for i in range(2):
            e.rollout(p, seed=i, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        for e, p in zip(engs, ins):
            e.rollout(p, seed=10 + i, **kw)
    torch.cuda.synchronize()
    print(f"{nsplit} engines serial: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per {B}-study rollout", flush=True)

    def worker(k, seed):
        with torch.cuda.stream(streams[k]):
            engs[k].rollout(ins[k], seed=seed, **kw)

    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            th = [threading.Thread(target=worker, args=(k, 20 + i)) for k in range(nsplit)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        torch.cuda.synchronize()
        print(f"{nsplit} engines concurrent: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per {B}-study rollout", flush=True)


if __name__ == "__main__":
    main()
