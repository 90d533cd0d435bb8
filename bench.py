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
`cpu_baseline`: the CPU oracle (port of the reference modules) on a bounded sample, all host cores: ONE study of the
           workload, every decode step of both rollouts executed (no extrapolation).
`gpu_eager_baseline`: the same oracle modules (plain functional PyTorch = what the reference's HF modules execute)
           run EAGERLY on this GPU under bf16 autocast on the same batch: the library-kernel bar.

Other BASELINE.json configs (each prints its own JSON line; not the driver's default):
    python bench.py --config 1     # single-tf, 1 image, batch 1, greedy 255 tokens
    python bench.py --config 2     # multi-tf, <= 5 images, batch 32, greedy 255 tokens
    python bench.py --config 3 [--all-params]   # longitudinal teacher-forced forward/backward, batch 64 (+ gradient all-reduce, N > 1)
    python bench.py --config 5     # generation sweep: 1 -> 512 studies per GPU, greedy 255 tokens
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
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[] index + 1")
    ap.add_argument("--all-params", action="store_true", help="config 3: gradients of every decoder parameter (default: LoRA only)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: DistributedSampler order instead of image-balanced shards")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline")
    ap.add_argument("--sweep-max", type=int, default=512)
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


# ----------------------------------------------------------------------------------------- synthetic batch
def global_image_counts(n_studies: int, max_images: int, seed: int = 1234):
    """valid images per study of the GLOBAL batch (all ranks), U{1..max_images}"""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(1, max_images + 1, (n_studies,), generator=g).tolist()


def make_study(a, gidx: int, n_img: int):
    """(pixels [N,3,384,384] with n_img valid images, prompt ids 1-D, label ids 1-D) of global study gidx"""
    import torch

    from cxrmate_b200 import synthetic as S
    px = S.make_images(1, a.images, seed=5000 + gidx, n_per_study=[n_img])[0]
    no_hist = gidx % 4 == 3
    prompt = S.make_prompts(1, a.prompt, seed=9000 + gidx, no_history_every=1 if no_hist else 0)[0]
    prompt = prompt[prompt != S.PAD]
    lab, lab_len = S.make_label_ids(1, 32, 256, seed=7000 + gidx)
    return px, prompt, lab[0, : int(lab_len[0])]


# ----------------------------------------------------------------------------------------- CPU arm
_CPU_W = {}


def cpu_study(a, cores: int, gidx: int):
    """ONE study of the workload through the oracle (fp32 port of the reference modules) on the host cores: CvT-21
    encode of its valid images (padded ones too, as the reference does), prompt pass + EVERY decode step of the sampled
    and the greedy rollout at batch 1, three CXR-BERT encodes (sample, greedy, label), cosine rewards.  Returns seconds."""
    import torch

    from cxrmate_b200 import synthetic as S
    from oracle import bert, cvt, decode, weights
    torch.set_num_threads(cores)
    if not _CPU_W:
        _CPU_W["sd"] = weights.make_cxrmate_weights(seed=0)
        _CPU_W["rsd"] = weights.make_cxrbert_weights(seed=1)
    sd, rsd = _CPU_W["sd"], _CPU_W["rsd"]
    counts = global_image_counts(max(a.studies, gidx + 1), a.images)
    px, prompt, lab = make_study(a, gidx, counts[gidx])
    T = a.tokens
    t = {}
    with torch.no_grad():
        t0 = time.perf_counter()
        mem, mask = cvt.encode_multi(sd, px[None])
        t["encode"] = time.perf_counter() - t0
        kw = dict(sections=S.SECTIONS, mask_token_id=S.PAD, eos_token_id=S.EOS, pad_token_id=S.PAD, max_new_tokens=T)
        t0 = time.perf_counter()
        smp = decode.rollout(sd, mem, mask, prompt[None], special_token_ids=S.SPECIAL_SAMPLE, do_sample=True, top_k=50,
                             generator=torch.Generator().manual_seed(gidx), **kw)
        grd = decode.rollout(sd, mem, mask, prompt[None], special_token_ids=S.SPECIAL_GREEDY, **kw)
        t["rollouts"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        m = S.id_map().long()
        embs = []
        for ids in (smp.sequences[0, len(prompt):], grd.sequences[0, len(prompt):]):
            w = m[ids[ids >= S.N_SPECIAL]]
            r = torch.cat((torch.tensor([S.RWD_CLS]), w, torch.tensor([S.RWD_SEP])))[None]
            embs.append(bert.cxrbert_cls_projection(rsd, r, torch.ones_like(r)))
        le = bert.cxrbert_cls_projection(rsd, lab[None], torch.ones_like(lab[None]))
        _ = [torch.nn.functional.cosine_similarity(e, le) for e in embs]
        t["reward"] = time.perf_counter() - t0
    total = sum(t.values())
    desc = (f"study {gidx}: {counts[gidx]} valid of {a.images} image slots (all slots encoded, as the reference does), prompt "
            f"{len(prompt)}, {smp.steps}+{grd.steps} decode steps executed at batch 1 (sample + greedy), 3 CXR-BERT encodes; "
            f"encode {t['encode']:.2f}s rollouts {t['rollouts']:.2f}s reward {t['reward']:.2f}s")
    return total, desc


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(a):
    """CPU arm: one step = ONE study of the workload (a bounded sample of the 32-study batch: the oracle at batch 32
    needs > 100 s per step on 16 cores), nothing extrapolated: `ms_per_step` is the measured time of that study and
    `value` = studies per second."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs, desc = [], ""
    for i in range(a.warmup + a.steps):
        sec, desc = cpu_study(a, cores, i % a.studies)
        if i >= a.warmup:
            secs.append(sec)
    per = sum(secs) / len(secs)
    v = 1.0 / per
    sample = (f"oracle (fp32 port of the reference modules), {cores} threads on {cpu_model()}: 1 study per step, cycling "
              f"through the studies of the workload batch; last: {desc}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * per, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "reference_step": "1 study (bounded sample of the 32-study step), every "
                   "decode step executed", "note": "CPU arm: the reference has no GPU kernels of its own; /root/reference "
                   "does not exist on the GPU box, so the oracle port of its modules is timed"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- eager PyTorch on the GPU
def gpu_eager_baseline(a, px_d, prompt_d, lab_d, lab_len_d, dev):
    """The oracle modules (functional PyTorch, the same ATen/cuBLAS/cuDNN kernels the reference's HF modules launch) on
    this GPU under bf16 autocast, same batch, same step: encode of all image slots -> sampled rollout -> greedy rollout
    (each its own KV-cached loop with torch.cat caches, as HF generate does) -> three CXR-BERT batches -> cosine.
    Device-timed with CUDA events; one untimed warm-up step of 8 tokens first."""
    import torch

    from cxrmate_b200 import synthetic as S
    from oracle import bert, cvt, decode, weights
    sd = {k: v.to(dev) for k, v in weights.make_cxrmate_weights(seed=0).items()}
    rsd = {k: v.to(dev) for k, v in weights.make_cxrbert_weights(seed=1).items()}
    m = S.id_map().long().to(dev)
    prompt = prompt_d.long()
    B = prompt.shape[0]

    def step(T):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            mem, mask = cvt.encode_multi(sd, px_d)
            kw = dict(sections=S.SECTIONS, mask_token_id=S.PAD, eos_token_id=S.EOS, pad_token_id=S.PAD, max_new_tokens=T)
            smp = decode.rollout(sd, mem, mask, prompt, special_token_ids=S.SPECIAL_SAMPLE, do_sample=True, top_k=50, **kw)
            grd = decode.rollout(sd, mem, mask, prompt, special_token_ids=S.SPECIAL_GREEDY, **kw)
            embs = []
            for seq in (smp.sequences[:, prompt.shape[1]:], grd.sequences[:, prompt.shape[1]:]):
                ids = torch.cat((torch.full((B, 1), S.RWD_CLS, device=dev), m[seq.clamp(min=S.N_SPECIAL)],
                                 torch.full((B, 1), S.RWD_SEP, device=dev)), 1)
                embs.append(bert.cxrbert_cls_projection(rsd, ids, torch.ones_like(ids)))
            lm = torch.arange(lab_d.shape[1], device=dev)[None] < lab_len_d[:, None]
            le = bert.cxrbert_cls_projection(rsd, lab_d.long(), lm)
            r = [torch.nn.functional.cosine_similarity(e.float(), le.float()) for e in embs]
            return (r[0] - r[1]).sum().item(), smp.steps + grd.steps

    step(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, nsteps = step(a.tokens)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    del sd, rsd
    torch.cuda.empty_cache()
    return {"value": B / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "dtype": "bf16 autocast", "steps": 1,
            "decode_steps_executed": nsteps,
            "what": f"oracle port of the reference modules, eager PyTorch {torch.__version__} on this GPU, batch {B}, all "
                    f"{a.images} image slots encoded (the reference encodes padded images), sample and greedy rollouts as "
                    "two KV-cached loops, device-timed, 1 step after an 8-token warm-up"}


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
                 rwd_max_seqs=3 * B, enc_chunk=64, use_cuda_graph=not a.no_graph)
    eng.load_state_dict(W.make_cxrmate_weights(seed=0))
    eng.load_state_dict(W.make_cxrbert_weights(seed=1), prefix="reward.")
    eng.finalize()
    eng.set_id_map(S.id_map(), S.RWD_CLS, S.RWD_SEP, S.BOS, S.SEP)

    # seeded synthetic inputs of the named shapes.  The GLOBAL batch is world x B studies; ranks
    # take B studies each: DistributedSampler order (--no-balance) or dealt so that every rank carries the same number
    # of valid IMAGES (cxrmate_b200.sharding.balance_by_images) - encoder and cross-attention cost follow images, and
    # the step time is the slowest rank's.
    from cxrmate_b200 import sharding
    counts = global_image_counts(world * B, N)
    if world > 1 and not a.no_balance:
        mine = sharding.balance_by_images(counts, world)[rank]
    else:
        mine = sharding.shard_studies(world * B, rank, world) if world > 1 else list(range(B))
    studies = [make_study(a, g, counts[g]) for g in mine]
    px_h = torch.stack([s_[0] for s_ in studies]).pin_memory()
    P_ = max(len(s_[1]) for s_ in studies)
    prompt_t = torch.full((B, P_), S.PAD, dtype=torch.int32)
    for b, s_ in enumerate(studies):
        prompt_t[b, : len(s_[1])] = s_[1].to(torch.int32)
    prompt_h = prompt_t.pin_memory()
    L_ = max(len(s_[2]) for s_ in studies)
    lab_t = torch.zeros(B, L_, dtype=torch.int32)
    for b, s_ in enumerate(studies):
        lab_t[b, : len(s_[2])] = s_[2].to(torch.int32)
    lab_h = lab_t.pin_memory()
    lab_len_h = torch.tensor([len(s_[2]) for s_ in studies], dtype=torch.int32).pin_memory()
    px_d, prompt_d, lab_d, lab_len_d = px_h.to(dev), prompt_h.to(dev), lab_h.to(dev), lab_len_h.to(dev)
    n_valid_images = int((px_h[:, :, 0, 0, 0] != 0).sum())
    prompt_tokens = [len(s_[1]) for s_ in studies]
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
        phases = eng.last_phase_ms()
        steps_exec = int(out["steps"].item())
        t_ms = torch.tensor([ms], device=dev)
        per_rank = torch.tensor([[ms / a.steps, float(n_valid_images)]], device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            allr = torch.empty(world, 2, device=dev)
            dist.all_gather_into_tensor(allr, per_rank)
            per_rank = allr
        ms = float(t_ms.item())
        per_rank_ms = [round(float(x), 3) for x in per_rank[:, 0]]
        per_rank_images = [int(x) for x in per_rank[:, 1]]
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

        # ---- the same step with the REAL text bridge in the loop (BPE decode + WordPiece encode on the host) ---------
        e2e_text = None
        if not a.no_e2e and rank == 0:
            try:
                from cxrmate_b200.scst import scst_step_text
                from cxrmate_b200.text_bridge import TextBridge
                dec_tok, rwd_tok = S.train_tokenizers()
                br = TextBridge(dec_tok, rwd_tok, S.BOS, S.SEP, S.EOS)
                lab_txt = [[rwd_tok.decode(s_[2].tolist(), skip_special_tokens=True)] for s_ in studies]
                tkw = dict(max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=S.PAD,
                           special_sample=S.SPECIAL_SAMPLE, sections_sample=S.SECTIONS[:3], special_greedy=S.SPECIAL_GREEDY,
                           sections_greedy=S.SECTIONS, top_k=50)
                scst_step_text(eng, br, px_h, prompt_h, lab_txt, seed=300, **tkw)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                bms = []
                for i in range(a.steps):
                    o_ = scst_step_text(eng, br, px_h, prompt_h, lab_txt, seed=301 + i, **tkw)
                    _ = o_["advantage"].cpu()
                    bms.append(o_["bridge_ms"])
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) / a.steps
                e2e_text = {"value": B / wall, "unit": UNIT, "ms_per_step": wall * 1000.0, "bridge_ms": sum(bms) / len(bms),
                            "n_gpus": 1,
                            "what": "cxrmate_b200.scst.scst_step_text on rank 0: pinned host pixels in, engine encode + rollouts, "
                                    "sequences to the host, split + byte-level BPE decode + WordPiece encode (30k / 30.5k "
                                    "vocabularies trained offline) on the host, engine CXR-BERT reward, advantage to the "
                                    "host; wall clock"}
            except Exception as ex:
                e2e_text = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

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
            # whole decode step: algorithmic bytes per step / device time per step of the timed (graph-replayed) run
            if roofline is not None and steps_exec > 1:
                esz = 2 if a.dtype == "bf16" else 4
                X = 6 * 2 * 768 * esz                                   # K and V bytes per cached token, all layers
                Wb = 73175040 * esz                                     # weights read once per step
                tok = n_valid_images * eng.tokens_per_image
                # self K/V: every row reads its visible prompt tokens + the tokens generated so far (mean over the steps)
                self_tok = 2 * sum(prompt_tokens) + 2 * B * (steps_exec - 1) / 2.0
                per_step = Wb + X * tok + X * self_tok + 2 * B * 30000 * 4
                dec_ms = (phases["rollout"] - phases["prompt_pass"]) / (steps_exec - 1)
                ach = per_step / (dec_ms / 1000.0) / 1e9
                roofline["decode_step"] = {
                    "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "algorithmic_bytes_per_step": int(per_step), "ms_per_step": dec_ms,
                    "bytes": "weights 146.4 MB + encoder K/V of the visible tokens (once per study) + self K/V of the visible "
                             "prompt and generated tokens of all 2B rows (mean over the steps) + fp32 logits",
                    "timing": "rollout phase minus prompt pass (CUDA events on the step's stream, graph replay), last timed step"}
            if a.profile_out:
                os.makedirs(os.path.dirname(os.path.abspath(a.profile_out)), exist_ok=True)
                json.dump({"breakdown": breakdown, "roofline": roofline}, open(a.profile_out, "w"), indent=1)

    eager = None
    if rank == 0 and world == 1 and not a.no_gpu_eager:
        try:
            eager = gpu_eager_baseline(a, px_d, prompt_d, lab_d, lab_len_d, dev)
        except Exception as ex:      # a baseline must never take the measured arm down
            eager = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sec, desc = cpu_study(a, cores, 0)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle (fp32 port of the reference modules), {cores} threads on {cpu_model()}: {desc}; nothing extrapolated"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "studies_per_gpu": B, "valid_images_rank0": n_valid_images,
                       "valid_images_per_rank": per_rank_images, "ms_per_step_per_rank": per_rank_ms,
                       "sharding": ("DistributedSampler order" if (a.no_balance or world == 1) else
                                    "studies dealt so that ranks carry equal valid-image counts (sharding.balance_by_images)"),
                       "prompt_len": int(prompt_h.shape[1]), "new_tokens": T, "decode_steps_executed": steps_exec,
                       "weights": "random-init cxrmate (CvT-21 + 6-layer BERT decoder + LoRA) and CXR-BERT-sized reward model",
                       "cache": "inputs larger than L2: 283 MB pixels and >1 GB of K/V per step, no explicit flush",
                       "cuda_graph": not a.no_graph},
            "decode_tokens_per_s": world * 2 * B * steps_exec * a.steps / (ms / 1000.0),
            "phase_ms": {k: round(v, 3) for k, v in phases.items()},
            "e2e": e2e, "e2e_text": e2e_text, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_eager_baseline": eager, "breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


# ----------------------------------------------------------------------------------------- configs 1, 2, 5: greedy generation
def run_generation(a):
    """BASELINE.json configs[0] (single-tf, 1 image, batch 1), configs[1] (multi-tf, <= 5 images, batch 32) and
    configs[4] (sweep 1 -> 512 studies per GPU): encode + cross K/V + greedy KV-cached decode of 255 tokens from [BOS]
    (special_token_ids=[SEP], default sections / positions: reference single.py:483-493, multi.py:218-228).
    One JSON line per batch size; `value` device-resident, `e2e` with pinned host pixels in and sequences out."""
    import torch

    from cxrmate_b200 import synthetic as S
    from cxrmate_b200 import synthetic_weights as W
    from cxrmate_b200.engine import Engine

    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    T = a.tokens
    N = 1 if a.config == 1 else a.images
    sizes = [1] if a.config == 1 else [a.studies] if a.config == 2 else \
        [b for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512) if b <= a.sweep_max]
    sd = {k: v for k, v in W.make_cxrmate_weights(seed=0).items() if "lora_" not in k}     # these variants carry no LoRA
    name = {1: "BASELINE.json configs[0] 'cxrmate-single-tf': 1 image 384x384, batch 1, greedy 255 tokens",
            2: f"BASELINE.json configs[1] 'cxrmate-multi-tf': <= {N} images/study, batch {a.studies}, greedy 255 tokens",
            5: f"BASELINE.json configs[4] generation sweep: B studies/GPU, <= {N} images/study, greedy 255 tokens"}[a.config]
    for B in sizes:
        line = {"metric": "generation_reports_per_sec", "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
                "config": {"workload": name, "studies_per_gpu": B}}
        eng = None
        try:
            eng = Engine(dtype=a.dtype, max_studies=B, max_images=N, max_prompt=8, max_new_tokens=T, rwd_layers=0,
                         enc_chunk=64, use_cuda_graph=not a.no_graph)
            eng.load_state_dict(sd)
            eng.finalize()
            counts = [1] * B if a.config == 1 else global_image_counts(B, N)
            px_h = torch.stack([S.make_images(1, N, seed=5000 + g, n_per_study=[counts[g]])[0] for g in range(B)]).pin_memory()
            px_d = px_h.to(dev)
            prompt = torch.full((B, 1), S.BOS, dtype=torch.int64, device=dev)
            kw = dict(mode="greedy", max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, mask_token_id=None,
                      special_greedy=[S.SEP], sections_greedy=[0, 1])

            def step(px):
                eng.encode(px)
                eng.prefill_cross_kv()
                return eng.rollout(prompt, **kw)

            for _ in range(a.warmup):
                out = step(px_d)
            torch.cuda.synchronize()
            l0 = eng.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                out = step(px_d)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            launches = eng.launch_count - l0
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(a.steps):
                o2 = step(px_h.to(dev, non_blocking=True))
                seq_h = o2.sequences.cpu()
            h1.record()
            torch.cuda.synchronize()
            ms2 = h0.elapsed_time(h1) / a.steps
            line.update({"value": B / (ms / 1000.0), "ms_per_step": ms, "decode_tokens_per_s": B * out.steps / (ms / 1000.0),
                         "e2e": {"value": B / (ms2 / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(px_h.numel() * 4),
                                 "d2h_bytes_per_step": int(seq_h.numel() * 8)},
                         "gpu_launches": int(launches)})
            line["config"].update({"valid_images": int(sum(counts)), "decode_steps_executed": out.steps,
                                   "workspace_gb": round(eng.workspace_bytes / 2 ** 30, 2)})
            if a.config == 1 and not a.no_cpu_baseline:      # the reference's own CPU-runnable case, same inputs
                from oracle import cvt, decode, weights
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                osd = {k: v for k, v in weights.make_cxrmate_weights(seed=0).items() if "lora_" not in k}
                with torch.no_grad():
                    t0 = time.perf_counter()
                    mem = cvt.encode_single(osd, px_h[:, 0])
                    o = decode.rollout(osd, mem, None, prompt.cpu(), special_token_ids=[S.SEP], sections=None,
                                       mask_token_id=None, max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD)
                    sec = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"the whole config on {cpu_model()}: 1 image encode + {o.steps} greedy steps, fp32, "
                                                  f"{sec:.2f}s"}
        except Exception as ex:
            line["error"] = f"{type(ex).__name__}: {ex}"[:400]
        finally:
            if eng is not None:
                eng.close()
            torch.cuda.empty_cache()
        print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- config 3: teacher-forced fwd/bwd
def run_teacher_forced(a):
    """BASELINE.json configs[2]: longitudinal teacher-forced forward/backward, batch 64 studies per GPU, prompt <= 256 +
    report 256 tokens (512 decoder positions): frozen CvT encode -> cross K/V -> decoder forward -> cross-entropy ->
    backward (cxrm_train_step; LoRA-only or every decoder parameter) -> gradient all-reduce over the ranks (NCCL),
    bucketed per decoder layer and overlapped with the backward pass."""
    import torch
    import torch.distributed as dist

    from cxrmate_b200 import synthetic as S
    from cxrmate_b200 import synthetic_weights as W
    from cxrmate_b200 import training
    from cxrmate_b200.engine import Engine
    from cxrmate_b200.modelling import position_ids_from_mask, token_ids_to_token_type_ids

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 64 if a.studies == 32 else a.studies
    N, L = a.images, 512
    eng = Engine(dtype=a.dtype, device=local, max_studies=B, max_images=N, max_prompt=8, max_new_tokens=8, rwd_layers=0,
                 enc_chunk=64, max_train_tokens=B * L)
    eng.load_state_dict(W.make_cxrmate_weights(seed=0))
    eng.finalize()
    counts = global_image_counts(world * B, N)
    px = torch.stack([S.make_images(1, N, seed=5000 + g, n_per_study=[counts[g]])[0] for g in range(rank * B, (rank + 1) * B)])
    px_h = px.pin_memory()
    g = torch.Generator().manual_seed(77 + rank)
    ids = torch.randint(S.N_SPECIAL, S.DEC_VOCAB, (B, L), generator=g)
    plen = torch.randint(8, 257, (B,), generator=g)            # prompt tokens, then [BOS] findings [SEP] impression [EOS]
    rlen = torch.randint(64, 257, (B,), generator=g)
    for b in range(B):
        p_, r_ = int(plen[b]), int(rlen[b])
        ids[b, 0], ids[b, p_ // 2], ids[b, p_ - 1] = S.PMT, S.PMT_SEP, S.BOS
        ids[b, p_ - 1 + r_ // 2] = S.SEP
        ids[b, p_ + r_ - 1:] = S.PAD
    labels = torch.roll(ids, -1, 1)
    labels[:, -1] = S.PAD
    for b in range(B):
        labels[b, : int(plen[b]) - 1] = S.PAD                  # the prompt is not a target
    mask = (ids != S.PAD).long()
    tt = token_ids_to_token_type_ids(ids, S.SPECIAL_GREEDY, S.SECTIONS)
    pos = position_ids_from_mask(mask)
    t_dev = [x.to(dev) for x in (ids, tt, pos, mask, labels)]
    lora = not a.all_params
    grads = torch.zeros(sum(ne for _, _, ne, _ in eng.grad_layout(lora)), dtype=torch.float32, device=dev)

    def step():
        eng.encode(px_h.to(dev, non_blocking=True))
        eng.prefill_cross_kv()
        return training.cross_entropy_backward(eng, *t_dev, pad_token_id=S.PAD, lora_only=lora, grads=grads, all_reduce=world > 1)

    for _ in range(a.warmup):
        loss, _ = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, _ = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / a.steps
    if rank == 0:
        n_tok = int((labels != S.PAD).sum())
        print(json.dumps({
            "metric": "teacher_forced_samples_per_sec", "value": world * B / (ms / 1000.0), "unit": "samples/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[2] 'cxrmate-tf longitudinal': batch 64/GPU, 512 decoder positions (prompt <= 256 + "
                                   "report <= 256), frozen encoder, teacher-forced forward + backward, " +
                                   ("LoRA-only gradients" if lora else "gradients of every decoder parameter"),
                       "counted_target_tokens_rank0": n_tok, "valid_images_rank0": int(sum(counts[rank * B:(rank + 1) * B])),
                       "gradient_bytes_all_reduced": int(grads.numel() * 4) if world > 1 else 0,
                       "all_reduce": "NCCL, one bucket per decoder layer, async on the NCCL stream under the backward" if world > 1 else "none"},
            "loss": float(loss.item()), "gpu_launches": int(eng.launch_count - l0),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == 3:
        run_teacher_forced(a)
    elif a.config != 4:
        run_generation(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
