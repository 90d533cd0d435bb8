"""Experiment: pipeline consecutive SCST steps -- run the image encoder (+ cross K/V projection) of batch n+1 on a second
stream while batch n's rollout (a latency-bound chain that leaves most SMs idle outside its attention kernels) decodes.

    python tools/exp_overlap_encode.py     # rollout alone | encode alone | both concurrently (wall, max of the two)
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
    counts = bench.global_image_counts(B, N)
    studies = [bench.make_study(a, g, counts[g]) for g in range(B)]
    px = torch.stack([s[0] for s in studies]).to(dev)
    P_ = max(len(s[1]) for s in studies)
    pr = torch.full((B, P_), S.PAD, dtype=torch.int64)
    for b, s in enumerate(studies):
        pr[b, : len(s[1])] = s[1]
    pr = pr.to(dev)
    sd = W.make_cxrmate_weights(seed=0)

    def mk():
        e = Engine(dtype="bf16", max_studies=B, max_images=N, max_prompt=a.prompt, max_new_tokens=T, rwd_layers=0, enc_chunk=64)
        e.load_state_dict(sd)
        e.finalize()
        return e

    kw = dict(mode="both", max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=S.PAD,
              special_sample=S.SPECIAL_SAMPLE, sections_sample=S.SECTIONS[:3], special_greedy=S.SPECIAL_GREEDY,
              sections_greedy=S.SECTIONS, top_k=50, temperature=1.0)
    dec, enc = mk(), mk()
    s_dec = torch.cuda.Stream(priority=int(os.environ.get("PRI_DEC", "-1")))
    s_enc = torch.cuda.Stream(priority=int(os.environ.get("PRI_ENC", "0")))
    with torch.cuda.stream(s_dec):
        dec.encode(px)
        dec.prefill_cross_kv()
        for i in range(2):
            dec.rollout(pr, seed=i, **kw)
    with torch.cuda.stream(s_enc):
        for i in range(2):
            enc.encode(px)
            enc.prefill_cross_kv()
    torch.cuda.synchronize()

    def run_dec(out):
        with torch.cuda.stream(s_dec):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec.rollout(pr, seed=7, **kw)
            e1.record()
        out["dec"] = (e0, e1)

    def run_enc(out, reps):
        with torch.cuda.stream(s_enc):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                enc.encode(px)
                enc.prefill_cross_kv()
            e1.record()
        out["enc"] = (e0, e1)

    def ms(p):
        return p[0].elapsed_time(p[1])

    for rep in range(2):
        o = {}
        run_dec(o)
        torch.cuda.synchronize()
        d_alone = ms(o["dec"])
        o = {}
        run_enc(o, 1)
        torch.cuda.synchronize()
        e_alone = ms(o["enc"])
        o = {}
        t0 = time.perf_counter()
        th = [threading.Thread(target=run_dec, args=(o,)), threading.Thread(target=run_enc, args=(o, 1))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        print(f"rollout alone {d_alone:.2f} ms | encode+xkv alone {e_alone:.2f} ms | concurrent: rollout {ms(o['dec']):.2f} ms, "
              f"encode+xkv {ms(o['enc']):.2f} ms, wall {wall:.2f} ms (serial sum {d_alone + e_alone:.2f})", flush=True)


if __name__ == "__main__":
    main()
