#!/usr/bin/env python
"""bench.py - SCST rollout throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0     # CPU arm (oracle port of the reference)

One "step" = one SCST rollout of `--studies` studies per GPU (BASELINE.json configs[3]): CvT-21 encode of the
valid images, cross-attention K/V, KV-cached sample (top-k 50) + greedy rollouts of `--tokens` new tokens,
CXR-BERT embeddings of the sample / greedy / label reports, cosine rewards, advantage.
Synthetic data and random-init weights of the named architecture (no network, no checkpoints).

`value`  : reports/s, inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`    : the same step through the C ABI with HOST (pinned) buffers: H2D of pixels/prompts/labels and D2H of
           sequences/log-probs/rewards inside the timed region (cxrm_scst_step_host; only the valid images of the
           padded pixel tensor cross PCIe, one encoder chunk at a time, overlapped with the encoding).
`roofline`: the dominant kernel (decode cross-attention over the encoder K/V cache), algorithmic bytes per launch /
           its mean launch duration from the engine's event profiler, vs the measured HBM copy peak.
`cpu_baseline`: the CPU oracle (port of the reference modules) on a bounded sample, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scst_reports_per_sec"
UNIT = "reports/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--studies", type=int, default=32, help="studies per GPU (weak scaling)")
    ap.add_argument("--images", type=int, default=5, help="max images per study")
    ap.add_argument("--prompt", type=int, default=256, help="max prompt length")
    ap.add_argument("--tokens", type=int, default=255, help="new tokens per rollout")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel-class event profile (no roofline)")
    ap.add_argument("--profile-out", default="")
    ap.add_argument("--cpu-tokens", type=int, default=12, help="decode steps of the bounded CPU sample")
    return ap.parse_args()


def workload_name(a):
    return (f"BASELINE.json configs[3] 'cxrmate SCST rollout': {a.studies} studies/GPU, 1..{a.images} images 384x384 "
            f"per study, prompt <= {a.prompt}, {a.tokens} new tokens, top-k-50 sample + greedy baseline + CXR-BERT "
            f"cosine reward")


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        # median over the samples taken under load (upper half: idle samples at the edges drag the median down)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_sample(a, cores: int):
    """Bounded sample of the SAME step on the host cores with the oracle (port of the reference modules,
    fp32): 1 study with 3 valid images, prompt 64, `cpu_tokens` decode steps of BOTH rollouts, three reward
    encodes of 256-token reports.  Scaled to the full step: the decode time is extrapolated linearly from
    cpu_tokens to --tokens steps (per-step cost grows slowly with the cache, so this flatters the CPU)."""
    import torch

    from cxrmate_b200 import synthetic as S
    from oracle import bert, cvt, decode, weights
    torch.set_num_threads(cores)
    sd = weights.make_cxrmate_weights(seed=0)
    rsd = weights.make_cxrbert_weights(seed=1)
    n_img = min(3, a.images)
    px = S.make_images(1, n_img, seed=1234, n_per_study=[n_img])
    Pc = min(64, a.prompt)
    g = torch.Generator().manual_seed(99)
    prompt = torch.cat((torch.tensor([S.PMT]), torch.randint(S.N_SPECIAL, S.DEC_VOCAB, (Pc - 12,), generator=g),
                        torch.tensor([S.PMT_SEP]), torch.randint(S.N_SPECIAL, S.DEC_VOCAB, (9,), generator=g),
                        torch.tensor([S.BOS])))[None]
    Ts = max(2, min(a.cpu_tokens, a.tokens))
    t = {}
    with torch.no_grad():
        t0 = time.perf_counter()
        mem, mask = cvt.encode_multi(sd, px)
        t["encode"] = time.perf_counter() - t0
        kw = dict(sections=S.SECTIONS, mask_token_id=S.PAD, eos_token_id=-1, pad_token_id=S.PAD)

        def both(n):
            t0 = time.perf_counter()
            decode.rollout(sd, mem, mask, prompt, special_token_ids=S.SPECIAL_SAMPLE, do_sample=True, top_k=50,
                           generator=torch.Generator().manual_seed(0), max_new_tokens=n, **kw)
            decode.rollout(sd, mem, mask, prompt, special_token_ids=S.SPECIAL_GREEDY, max_new_tokens=n, **kw)
            return time.perf_counter() - t0

        t["prefill"] = both(1)                    # prompt pass + first token of both rollouts
        t["decode_sample"] = both(Ts)
        ids, lens = S.make_label_ids(3, 256, 256, seed=7)
        m = torch.arange(ids.shape[1])[None] < lens[:, None]
        t0 = time.perf_counter()
        bert.cxrbert_cls_projection(rsd, ids, m)
        t["reward"] = time.perf_counter() - t0
    per_step = (t["decode_sample"] - t["prefill"]) / (Ts - 1)
    per_study = t["encode"] + t["prefill"] + per_step * (a.tokens - 1) + t["reward"]
    sample = (f"oracle (port of the reference modules), fp32, {cores} threads: 1 study, {n_img} images 384x384, "
              f"prompt {prompt.shape[1]}, {Ts} decode steps x (sample+greedy) extrapolated linearly to {a.tokens} "
              f"({per_step * 1000:.0f} ms per step pair), 3 x 256-token CXR-BERT encodes; measured encode "
              f"{t['encode']:.2f}s prefill {t['prefill']:.2f}s decode({Ts}) {t['decode_sample']:.2f}s "
              f"reward {t['reward']:.2f}s")
    return 1.0 / per_study, sample, per_study


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(a.warmup + a.steps):
        v, sample, per = cpu_sample(a, cores)
        if i >= a.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * a.studies / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU arm: the reference has no GPU kernels of its own; "
                   "/root/reference does not exist on the GPU box, so the oracle port of its modules is timed"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    from cxrmate_b200 import synthetic as S
    from cxrmate_b200 import synthetic_weights as W
    from cxrmate_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, N, T = a.studies, a.images, a.tokens

    eng = Engine(dtype=a.dtype, device=local, max_studies=B, max_images=N, max_prompt=a.prompt, max_new_tokens=T,
                 rwd_max_seqs=3 * B, enc_chunk=32, use_cuda_graph=not a.no_graph)
    eng.load_state_dict(W.make_cxrmate_weights(seed=0))
    eng.load_state_dict(W.make_cxrbert_weights(seed=1), prefix="reward.")
    eng.finalize()
    eng.set_id_map(S.id_map(), S.RWD_CLS, S.RWD_SEP, S.BOS, S.SEP)

    # seeded synthetic inputs of the named shapes, different per rank
    px_h = S.make_images(B, N, seed=1234 + rank).pin_memory()
    prompt_h = S.make_prompts(B, a.prompt, seed=99 + rank).to(torch.int32).pin_memory()
    lab, lab_len = S.make_label_ids(B, 32, 256, seed=7 + rank)
    lab_h, lab_len_h = lab.to(torch.int32).pin_memory(), lab_len.to(torch.int32).pin_memory()
    px_d, prompt_d, lab_d, lab_len_d = px_h.to(dev), prompt_h.to(dev), lab_h.to(dev), lab_len_h.to(dev)
    n_valid_images = int((px_h[:, :, 0, 0, 0] != 0).sum())
    kw = dict(max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=S.PAD,
              special_sample=S.SPECIAL_SAMPLE, sections_sample=S.SECTIONS[:3], special_greedy=S.SPECIAL_GREEDY,
              sections_greedy=S.SECTIONS, top_k=50, temperature=1.0)
    stream = torch.cuda.Stream(device=dev)
    gather_buf = torch.empty(world, 2, B, device=dev) if world > 1 else None

    def step_device(i):
        out = eng.scst_step(px_d, prompt_d, lab_d, lab_len_d, seed=i, **kw)
        if world > 1:   # the only collective of the path: gather per-rank rewards and baselines
            dist.all_gather_into_tensor(gather_buf, torch.stack((out["reward"], out["baseline"])))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for i in range(a.warmup):
            out = step_device(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(a.steps):
            out = step_device(a.warmup + i)
        ev1.record(stream)
        barrier()
        launches = eng.launch_count - l0
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        steps_exec = int(out["steps"].item())
        t_ms = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms.item())
        value = world * B * a.steps / (ms / 1000.0)

        # ---- end to end through the C ABI with host buffers ------------------------------------------
        e2e = None
        if not a.no_e2e:
            outs = None
            for i in range(1):
                outs = eng.scst_step(px_h, prompt_h, lab_h, lab_len_h, seed=100 + i, **kw)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(a.steps):
                outs = eng.scst_step(px_h, prompt_h, lab_h, lab_len_h, seed=200 + i, out=outs, **kw)
            e1.record(stream)
            barrier()
            t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            # the engine copies only the valid (non-padding) images of the host pixel tensor, chunk by chunk
            h2d = (n_valid_images * px_h[0, 0].numel() * 4 + B * N + prompt_h.numel() * 4 + lab_h.numel() * 4 +
                   lab_len_h.numel() * 4)
            d2h = sum(v.numel() * v.element_size() for v in outs.values())
            e2e = {"value": world * B * a.steps / (float(t2.item()) / 1000.0), "unit": UNIT,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

        # ---- per-kernel-class profile of one more step (event pairs around every launch, no graph) ------
        roofline, breakdown = None, None
        if rank == 0 and not a.no_profile:
            eng.set_profile(True)
            eng.scst_step(px_d, prompt_d, lab_d, lab_len_d, seed=999, **kw)
            rep = eng.profile_report()
            eng.set_profile(False)
            total = sum(v["ms"] for v in rep.values())
            breakdown = {k: {"ms": round(v["ms"], 3), "n": v["n"], "share": round(v["ms"] / total, 4)}
                         for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
            ca = rep.get("decode.cross_attn")
            if ca:
                tok = n_valid_images * eng.tokens_per_image            # visible encoder tokens of this rank's studies
                esz = 2 if a.dtype == "bf16" else 4
                bytes_per_launch = tok * 2 * 768 * esz                  # K and V rows of ONE layer, read once per study
                dur = ca["ms"] / ca["n"] / 1000.0
                ach = bytes_per_launch / dur / 1e9
                traffic = None
                try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture
                    tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
                    if tj.get("valid_images") == n_valid_images and tj.get("dtype") == a.dtype:
                        traffic = tj["bytes_per_launch"]
                except Exception:
                    pass
                roofline = {"kernel": "decode_cross_persist_kernel<2> (decode-step cross-attention of ONE layer: all studies, "
                                      "sample+greedy rows share each K/V read)",
                            "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                            "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": bytes_per_launch,
                            "timing": "CUDA event pair around each launch on the launching stream (engine profiler, graph "
                                      "bypassed; includes ~3 us of launch/event overhead per launch)",
                            "mean_launch_us": dur * 1e6, "launches_profiled": ca["n"],
                            "share_of_step": round(ca["ms"] / total, 4)}
            if a.profile_out:
                os.makedirs(os.path.dirname(os.path.abspath(a.profile_out)), exist_ok=True)
                json.dump({"breakdown": breakdown, "roofline": roofline}, open(a.profile_out, "w"), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample, _ = cpu_sample(a, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "studies_per_gpu": B, "valid_images_rank0": n_valid_images,
                       "prompt_len": int(prompt_h.shape[1]), "new_tokens": T, "decode_steps_executed": steps_exec,
                       "weights": "random-init cxrmate (CvT-21 + 6-layer BERT decoder + LoRA) and CXR-BERT-sized reward model",
                       "cache": "inputs larger than L2: 283 MB pixels and >1 GB of K/V per step, no explicit flush",
                       "cuda_graph": not a.no_graph},
            "decode_tokens_per_s": world * 2 * B * steps_exec * a.steps / (ms / 1000.0),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
