"""The decode contract, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's `generate()` is HF `GenerationMixin._sample` driven by the
reference's `prepare_inputs_for_generation`
(modules/transformers/longitudinal_model/modelling_longitudinal.py:251-295),
written for transformers 4.36/4.41.  Under the installed transformers 5.5.0
that override mis-feeds the cache (SURVEY.md section 0 finding 4), so this
file restates the 4.41 semantics directly:

* token types  - token_ids_to_token_type_ids       (:297-338)
                 token_ids_to_token_type_ids_past  (:340-364)
* mask/positions - `ids != mask_token_id`, `relu(cumsum(mask) - 1)` (:274-277, :283)
* heads        - greedy argmax; TopKLogitsWarper (HF generation/logits_process.py
                 :536-587) -> softmax -> multinomial(1) which ATen evaluates as
                 argmax(p / q), q ~ Exp(1) (aten/src/ATen/native/Distributions.cpp
                 fast path); EOS bookkeeping (HF generation/utils.py:2788-2805).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import bert


def token_type_ids_full(ids: torch.Tensor, special_token_ids, sections=None) -> torch.Tensor:
    """Vectorised equivalent of the reference's double loop (:297-338): for the
    i-th special token, columns strictly after its FIRST occurrence get
    sections[i+1], unless it is absent, sits at column 0 (argmax()==0 is the
    reference's "absent" test) or is the last column."""
    sections = sections if sections is not None else list(range(len(special_token_ids) + 1))
    B, L = ids.shape
    tt = torch.full_like(ids, sections[0])
    col_idx = torch.arange(L, device=ids.device)[None]
    for i, tok in enumerate(special_token_ids):
        cols = (ids == tok).int().argmax(dim=1) + 1
        ok = (cols != 1) & (cols < L)
        sel = ok[:, None] & (col_idx >= cols[:, None])
        tt = torch.where(sel, torch.full_like(tt, sections[i + 1]), tt)
    return tt


def token_type_ids_past(ids: torch.Tensor, special_token_ids, sections=None) -> torch.Tensor:
    """(:340-364) type of the LAST token of `ids`: sections[i+1] of the last
    listed special token that occurs anywhere in ids[:, :-1]."""
    sections = sections if sections is not None else list(range(len(special_token_ids) + 1))
    tt = torch.full((ids.shape[0], 1), sections[0], dtype=torch.long, device=ids.device)
    prev = ids[:, :-1]
    for i, tok in enumerate(special_token_ids):
        exists = torch.any(prev == tok, dim=1, keepdim=True)
        tt[exists] = sections[i + 1]
    return tt


def positions_from_mask(mask: torch.Tensor) -> torch.Tensor:
    return torch.relu(torch.cumsum(mask.to(torch.int64), dim=1) - 1)


def top_k_mask(scores: torch.Tensor, k: int) -> torch.Tensor:
    """TopKLogitsWarper: everything strictly below the k-th largest -> -inf (ties kept)."""
    k = min(k, scores.shape[-1])
    thr = torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(scores < thr, float("-inf"))


@dataclass
class RolloutResult:
    sequences: torch.Tensor          # [B, P + steps] int64 (no auto-prepended BOS)
    steps: int
    scores: list                     # steps x [B,V] fp32: raw logits (greedy) or top-k-masked (sample)
    logprobs: torch.Tensor           # [B, steps] fp32 log-prob of the emitted token (0 at PAD fill)
    margins: torch.Tensor            # [B, steps] top1 - top2 of the decision variable (tie diagnosis)


def rollout(sd, memory, memory_mask, prompt_ids, *, special_token_ids, sections, mask_token_id, max_new_tokens,
            eos_token_id, pad_token_id, do_sample=False, top_k=50, temperature=1.0, exp_noise=None,
            generator=None, layers=6, use_cache=True) -> RolloutResult:
    """Greedy or top-k multinomial decode of the reference models.

    mask_token_id None  -> multi/single variants: mask all ones, positions arange
                           (modelling_multi.py:229-261).
    exp_noise           -> optional [max_new_tokens, B, V] fp32 Exp(1) draws; when
                           absent they are drawn step by step from `generator`
                           exactly as torch.multinomial would.
    use_cache=False     -> recompute the full sequence every step (the
                           property "cached == uncached" is a test).
    """
    ids = prompt_ids.clone()
    B = ids.shape[0]
    unfinished = torch.ones(B, dtype=torch.bool, device=ids.device)
    cache = bert.DecoderCache() if use_cache else None
    scores, lps, margins = [], [], []
    steps = 0
    for t in range(max_new_tokens):
        mask = (ids != mask_token_id).to(torch.int64) if mask_token_id is not None else torch.ones_like(ids)
        pos = positions_from_mask(mask)
        if use_cache and t > 0:
            feed = ids[:, -1:]
            tt = token_type_ids_past(ids, special_token_ids, sections)
            pos_in = pos[:, -1:]
        else:
            feed = ids
            tt = token_type_ids_full(ids, special_token_ids, sections)
            pos_in = pos
        if cache is not None and not use_cache:
            cache = None
        logits = bert.decoder_logits(sd, feed, tt, pos_in, mask, memory, memory_mask, cache, layers, last_only=True)
        logits = logits[:, -1].float()
        if do_sample:
            s = logits / temperature if temperature != 1.0 else logits
            s = top_k_mask(s, top_k)
            p = torch.softmax(s, dim=-1)
            if exp_noise is not None:
                q = exp_noise[t]
            else:
                q = torch.empty_like(p).exponential_(1, generator=generator)
            r = p / q
            nxt = torch.argmax(r, dim=-1)
            top2 = torch.topk(r, 2)[0]
            margins.append((top2[:, 0] - top2[:, 1]) / top2[:, 0])
            lp = torch.log_softmax(s, dim=-1).gather(1, nxt[:, None])[:, 0]
        else:
            s = logits
            nxt = torch.argmax(s, dim=-1)
            top2 = torch.topk(s, 2)[0]
            margins.append(top2[:, 0] - top2[:, 1])
            lp = torch.log_softmax(s, dim=-1).gather(1, nxt[:, None])[:, 0]
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_token_id))
        lp = torch.where(nxt != pad_token_id, lp, torch.zeros_like(lp))
        scores.append(s)
        lps.append(lp)
        ids = torch.cat((ids, nxt[:, None]), dim=1)
        unfinished = unfinished & (nxt != eos_token_id)
        steps += 1
        if not bool(unfinished.any()):
            break
    return RolloutResult(ids, steps, scores, torch.stack(lps, 1), torch.stack(margins, 1))
