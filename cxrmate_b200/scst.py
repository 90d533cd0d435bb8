"""`scst_step` of the reference with REAL strings in the loop, driven through the engine
(reference modules/lightning_modules/longitudinal/scst/gen_prompt.py:174-259; scst/gt_prompt.py:62-142).

    encode once -> shared cross K/V -> sample + greedy rollouts in one decode loop -> [CPU] split / BPE decode /
    WordPiece encode -> CXR-BERT embeddings of sample, greedy and label reports in one batch -> cosine rewards ->
    advantage = sample - baseline (:241) -> (optionally) REINFORCE backward of the sampled rows.

`cxrm_scst_step_host` (the benchmarked call) replaces the CPU stage by a device-side id map because no vocabularies ship
with the repository; this module is the variant a reference user switches to: it takes the decoder tokenizer and the
CXR-BERT tokenizer and returns the strings the reference logs.
"""
from __future__ import annotations

import time
from typing import List, Optional

import torch

from .engine import Engine
from .text_bridge import TextBridge


def scst_step_text(engine: Engine, bridge: TextBridge, pixels: torch.Tensor, prompt_ids: torch.Tensor,
                   labels: List[List[str]], *, max_new_tokens: int, eos_token_id: int, pad_token_id: int, mask_token_id: int,
                   special_sample, sections_sample, special_greedy, sections_greedy, top_k: int = 50,
                   temperature: float = 1.0, seed: int = 0, exp_noise: Optional[torch.Tensor] = None):
    """pixels [B,N,3,H,W] fp32 (host pinned or device), prompt_ids [B,P] int (same place), labels: list of lists of
    strings (one report per study, as CXRBERTReward takes them).  Returns a dict with device tensors `sequences`
    [2B, P+T] (sample rows first), `logprobs`, `reward`, `baseline`, `advantage`, the strings `sample_str`,
    `baseline_str`, and `bridge_ms` (host time of the text stage)."""
    dev = engine.device
    B = pixels.shape[0]
    lab_future = bridge.encode_async([j for i in labels for j in i], key="label")     # CPU, under the GPU work below
    px = pixels if pixels.is_cuda else pixels.to(dev, non_blocking=True)
    pr = prompt_ids if prompt_ids.is_cuda else prompt_ids.to(dev, non_blocking=True)
    engine.encode(px)
    engine.prefill_cross_kv()
    out = engine.rollout(pr.long(), mode="both", max_new_tokens=max_new_tokens, eos_token_id=eos_token_id,
                         pad_token_id=pad_token_id, mask_token_id=mask_token_id, special_sample=special_sample,
                         sections_sample=sections_sample, special_greedy=special_greedy, sections_greedy=sections_greedy,
                         top_k=top_k, temperature=temperature, exp_noise=exp_noise, seed=seed)
    P = pr.shape[1]
    seq = out.sequences[:, : P + out.steps]
    t0 = time.perf_counter()
    texts, ids_h, lens_h = bridge(seq)                        # D2H + split + BPE decode + WordPiece encode
    lab_ids_h, lab_lens_h = lab_future.result()
    bridge_ms = (time.perf_counter() - t0) * 1000.0
    # one reward batch: sample rows, greedy rows, labels, padded to the longest of the three
    Lr = max(ids_h.shape[1], lab_ids_h.shape[1])
    allids = torch.zeros(3 * B, Lr, dtype=torch.int32, device=dev)
    allids[: 2 * B, : ids_h.shape[1]].copy_(ids_h, non_blocking=True)
    allids[2 * B:, : lab_ids_h.shape[1]].copy_(lab_ids_h, non_blocking=True)
    alllens = torch.cat((lens_h.to(dev, non_blocking=True), lab_lens_h.to(dev, non_blocking=True)))
    cap = max(1, (engine.cfg.rwd_max_seqs * engine.cfg.rwd_max_len) // Lr)
    emb = torch.cat([engine.reward_embed(allids[i:i + cap], alllens[i:i + cap]) for i in range(0, 3 * B, cap)])
    reward = engine.cosine(emb[:B], emb[2 * B:])
    baseline = engine.cosine(emb[B:2 * B], emb[2 * B:])
    return dict(sequences=seq, logprobs=out.logprobs[:, : out.steps], steps=out.steps, reward=reward, baseline=baseline,
                advantage=reward - baseline, sample_str=texts[:B], baseline_str=texts[B:], bridge_ms=bridge_ms, rollout=out)
