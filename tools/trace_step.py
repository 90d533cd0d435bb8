"""In-graph timeline of the decode step: run a short rollout under torch.profiler (CUPTI activity records carry the
start / end of every kernel, CUDA-graph nodes included) and print, per kernel of one mid-rollout step, start offset,
duration (which includes the time spent in griddepcontrol.wait) and the increment of the chain's end time.   python tools/trace_step.py [tokens]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from cxrmate_b200 import synthetic as S  # noqa: E402
from cxrmate_b200 import synthetic_weights as W  # noqa: E402
from cxrmate_b200.engine import Engine  # noqa: E402


class A:
    studies, images, prompt, tokens = 32, 5, 256, 255


def main():
    a = A()
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    dev = torch.device("cuda", 0)
    B, N = a.studies, a.images
    counts = bench.global_image_counts(B, N)
    studies = [bench.make_study(a, g, counts[g]) for g in range(B)]
    px = torch.stack([s[0] for s in studies]).to(dev)
    P_ = max(len(s[1]) for s in studies)
    pr = torch.full((B, P_), S.PAD, dtype=torch.int64)
    for b, s in enumerate(studies):
        pr[b, : len(s[1])] = s[1]
    pr = pr.to(dev)
    e = Engine(dtype="bf16", max_studies=B, max_images=N, max_prompt=a.prompt, max_new_tokens=a.tokens, rwd_layers=0, enc_chunk=64)
    e.load_state_dict(W.make_cxrmate_weights(seed=0))
    e.finalize()
    kw = dict(mode="both", max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=S.PAD,
              special_sample=S.SPECIAL_SAMPLE, sections_sample=S.SECTIONS[:3], special_greedy=S.SPECIAL_GREEDY,
              sections_greedy=S.SECTIONS, top_k=50, temperature=1.0)
    e.encode(px)
    e.prefill_cross_kv()
    for i in range(2):
        e.rollout(pr, seed=i, want_margins=False, **kw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e.rollout(pr, seed=5, want_margins=False, **kw)
        torch.cuda.synchronize()
    ev = [(x.name, x.time_range.start, x.time_range.end) for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda t: t[1])
    # a step ends with the sampling kernel (which also embeds the next step's input); the first one closes the prompt pass
    ends = [i for i, x in enumerate(ev) if "sample_step" in x[0]]
    print(f"{len(ev)} kernel records, {len(ends)} sampling kernels")
    if len(ends) < 5:
        for x in ev[:50]:
            print(x)
        return
    k = len(ends) - 3            # a late step (longest caches)
    seg = ev[ends[k - 1] + 1: ends[k] + 1]
    nxt = ev[ends[k] + 1][1] if ends[k] + 1 < len(ev) else seg[-1][2]
    t0 = seg[0][1]
    prev_end = t0
    rows = []
    for name, s_, e_ in seg:
        short = name.replace("void ", "").replace("cxrm::", "").replace("(anonymous namespace)::", "").split("(")[0][:44]
        rows.append((short, s_ - t0, e_ - s_, e_ - prev_end))
        prev_end = max(prev_end, e_)
    print(f"step {k}: {len(seg)} kernels, {prev_end - t0:.1f} us from the first start to the last end; next step starts {nxt - t0:.1f} us after this one")
    print(f"{'kernel':46s} {'start':>8s} {'dur':>7s} {'end+':>7s}")
    for r in rows:
        print(f"{r[0]:46s} {r[1]:8.1f} {r[2]:7.1f} {r[3]:7.1f}")
    agg = {}
    for r in rows:
        a_ = agg.setdefault(r[0], [0, 0.0, 0.0])
        a_[0] += 1
        a_[1] += r[2]
        a_[2] += max(r[3], 0.0)
    print("\nper kernel: n, sum of durations, sum of end-time increments (= its share of the step's critical path)")
    for kname, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{kname:46s} {v[0]:3d} {v[1]:8.1f} {v[2]:8.1f}")
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "step_trace.json"), "w"))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
