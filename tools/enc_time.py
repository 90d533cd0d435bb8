"""Encoder timing: GPU time (CUDA events) vs host enqueue time of cxrm_encode for the benchmark batch (100 valid images).
    python tools/enc_time.py [n_studies]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cxrmate_b200 import synthetic_weights as W  # noqa: E402
from cxrmate_b200.engine import Engine  # noqa: E402


class A:
    studies, images, prompt, tokens = 32, 5, 256, 255


def main():
    a = A()
    counts = bench.global_image_counts(a.studies, a.images)
    px = torch.stack([bench.make_study(a, g, counts[g])[0] for g in range(a.studies)]).cuda()
    e = Engine(dtype="bf16", max_studies=a.studies, max_images=a.images, max_prompt=a.prompt, max_new_tokens=a.tokens, rwd_layers=0,
               enc_chunk=int(os.environ.get("ENC_CHUNK", "64")))
    e.load_state_dict(W.make_cxrmate_weights(seed=0))
    e.finalize()
    reps = int(os.environ.get("ENC_REPS", "3"))
    for _ in range(reps):
        e.encode(px)
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        e.encode(px)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"encode {sum(counts)} images: GPU {e0.elapsed_time(e1):.2f} ms, host enqueue {(t1 - t0) * 1e3:.2f} ms, launches {e.launch_count}")
    if os.environ.get("PROFILE"):
        e.set_profile(True)
        e.encode(px)
        rep = e.profile_report()
        e.set_profile(False)
        tot = sum(v["ms"] for v in rep.values())
        for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"  {k:28s} {v['ms']:8.3f} ms  n={v['n']:4d}  {v['ms'] / v['n'] * 1e3:8.1f} us/launch")
        print(f"  total {tot:.2f} ms")


if __name__ == "__main__":
    main()
