"""Beam search (reference test_step: `generate(num_beams=self.num_test_beams)`), restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference call sites: modules/lightning_modules/longitudinal/gt_prompt.py:344-362, gen_prompt.py:184-200,
modules/lightning_modules/single.py:552-562, multi.py:265-275 (num_beams = num_test_beams, default flags otherwise:
length_penalty 1.0, early_stopping False, num_return_sequences 1, no logits processors).

The algorithm is HF `GenerationMixin._beam_search` of the installed transformers 5.5.0
($SP/transformers/generation/utils.py:3076-3385 with the helpers :2878-3073), which its refactor states reproduces the
4.x `BeamSearchScorer` results under the default flags.  Per step, for every study:

  1. log_softmax of the fp32 logits of each running beam + that beam's accumulated score           (:3259-3281)
  2. the 2*num_beams best continuations over the flattened (beam, token) axis                      (:2948-2996)
  3. a continuation "hits" when its token is EOS or the sequence reaches max_length                 (:3300-3306)
  4. next running beams = the num_beams best continuations that did not hit (hits are pushed down by -1e9)  (:2998-3017)
  5. finished set = best num_beams of {old finished} U {hits among the first num_beams continuations}, scored
     sum_logprob / generated_len**length_penalty; nothing is added once the early-stop heuristic is satisfied  (:3019-3073)
  6. heuristic (early_stopping False): the study can still improve iff some finished slot is empty or
     best_running_score / generated_len**length_penalty > worst finished score                     (:2878-2921)
  7. stop when no study can improve, or every continuation of every study hit (max length)          (:2923-2946)

The model is a callback `step_fn(flat_ids [B*nb, len], beam_idx | None) -> logits [B*nb, V]` (row = study * nb + beam,
HF's layout); `beam_idx` [B*nb] is the cache reorder to apply BEFORE evaluating the step (None on the first call, which
sees the whole prompt).  tests/test_oracle.py pins this loop against `transformers` `generate(num_beams=...)` on a
small HF decoder.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import bert
from .decode import positions_from_mask, token_type_ids_full, token_type_ids_past

NEG = -1.0e9


@dataclass
class BeamResult:
    sequences: torch.Tensor          # [B, P + max generated length] best finished hypothesis, PAD filled
    scores: torch.Tensor             # [B] its length-penalised score
    steps: int                       # decode steps executed
    all_sequences: torch.Tensor      # [B, nb, P + T] the whole finished set (diagnostics)
    all_scores: torch.Tensor         # [B, nb]


def _take(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """t [B, n, ...], idx [B, m] -> [B, m, ...]"""
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


def beam_search(step_fn, prompt_ids: torch.Tensor, *, num_beams: int, max_new_tokens: int, eos_token_id: int,
                pad_token_id: int, length_penalty: float = 1.0) -> BeamResult:
    B, P = prompt_ids.shape
    nb, K, T = num_beams, 2 * num_beams, max_new_tokens
    dev = prompt_ids.device
    max_len = P + T
    fill = pad_token_id or eos_token_id      # HF's `output_fill_value = pad_token_id or eos_token_id[0]` (:3181): pad id 0 -> EOS
    run_seq = torch.full((B, nb, max_len), fill, dtype=torch.int64, device=dev)
    run_seq[:, :, :P] = prompt_ids[:, None, :]
    fin_seq = run_seq.clone()
    run_score = torch.zeros(B, nb, device=dev)
    run_score[:, 1:] = NEG
    fin_score = torch.full((B, nb), NEG, device=dev)
    fin_len = torch.zeros(B, nb, dtype=torch.int64, device=dev)
    is_fin = torch.zeros(B, nb, dtype=torch.bool, device=dev)
    can_improve = torch.ones(B, 1, dtype=torch.bool, device=dev)
    top_nb = (torch.arange(K, device=dev) < nb)[None]
    beam_idx = None
    cur = P
    steps = 0
    while True:
        logits = step_fn(run_seq[:, :, :cur].reshape(B * nb, cur), beam_idx).float()
        V = logits.shape[-1]
        acc = (torch.log_softmax(logits, dim=-1).view(B, nb, V) + run_score[:, :, None]).view(B, nb * V)
        cand_score, flat = torch.topk(acc, K, dim=1)                       # sorted, best first
        src, tok = flat // V, flat % V
        cand_seq = _take(run_seq, src)
        cand_seq[:, :, cur] = tok
        hit = (tok == eos_token_id) | (cur + 1 >= max_len)
        # next running beams
        keep = torch.topk(cand_score + hit.float() * NEG, nb, dim=1)[1]
        run_seq, run_score = _take(cand_seq, keep), _take(cand_score + hit.float() * NEG, keep)
        beam_idx = (_take(src, keep) + torch.arange(B, device=dev)[:, None] * nb).reshape(-1)
        # finished set
        just = hit & top_nb
        s = cand_score / float(cur + 1 - P) ** length_penalty
        s = s + (~can_improve).float() * NEG + (~just).float() * NEG
        m_seq, m_score = torch.cat((fin_seq, cand_seq), 1), torch.cat((fin_score, s), 1)
        m_fin = torch.cat((is_fin, just), 1)
        m_len = torch.cat((fin_len, torch.full((B, K), cur + 1 - P, dtype=torch.int64, device=dev)), 1)
        best = torch.topk(m_score, nb, dim=1)[1]
        fin_seq, fin_score, is_fin, fin_len = _take(m_seq, best), _take(m_score, best), _take(m_fin, best), _take(m_len, best)
        cur += 1
        steps += 1
        # early-stop heuristic (early_stopping False)
        best_running = run_score[:, :1] / float(cur - P) ** length_penalty
        worst_fin = torch.where(is_fin, fin_score.min(dim=1, keepdim=True)[0], torch.full_like(fin_score, NEG))
        can_improve = can_improve & torch.any(best_running > worst_fin, dim=-1, keepdim=True)
        if not (bool(can_improve.any()) and not bool(hit.all())):
            break
    glen = torch.where(is_fin[:, 0], fin_len[:, 0], torch.zeros_like(fin_len[:, 0]))
    out_len = P + int(glen.max())
    return BeamResult(fin_seq[:, 0, :out_len], fin_score[:, 0], steps, fin_seq, fin_score)


def cxrmate_step_fn(sd, memory, memory_mask, *, num_beams: int, special_token_ids, sections, mask_token_id, layers=6):
    """step_fn over the oracle's decoder (oracle/bert.py) with the reference's mask / position / token-type rules
    (oracle/decode.py) and a KV cache that is reordered by `beam_idx` like `Cache.reorder_cache` does."""
    mem = memory.repeat_interleave(num_beams, dim=0)
    mmask = None if memory_mask is None else memory_mask.repeat_interleave(num_beams, dim=0)
    cache = bert.DecoderCache()

    def step(ids, beam_idx):
        mask = (ids != mask_token_id).to(torch.int64) if mask_token_id is not None else torch.ones_like(ids)
        pos = positions_from_mask(mask)
        if beam_idx is None:
            feed, tt, pos_in = ids, token_type_ids_full(ids, special_token_ids, sections), pos
        else:
            cache.self_k = [k.index_select(0, beam_idx) for k in cache.self_k]
            cache.self_v = [v.index_select(0, beam_idx) for v in cache.self_v]
            feed, tt, pos_in = ids[:, -1:], token_type_ids_past(ids, special_token_ids, sections), pos[:, -1:]
        return bert.decoder_logits(sd, feed, tt, pos_in, mask, mem, mmask, cache, layers, last_only=True)[:, -1]

    return step


def beam_rollout(sd, memory, memory_mask, prompt_ids, *, num_beams, special_token_ids, sections, mask_token_id,
                 max_new_tokens, eos_token_id, pad_token_id, length_penalty=1.0, layers=6) -> BeamResult:
    step = cxrmate_step_fn(sd, memory, memory_mask, num_beams=num_beams, special_token_ids=special_token_ids,
                           sections=sections, mask_token_id=mask_token_id, layers=layers)
    with torch.no_grad():
        return beam_search(step, prompt_ids, num_beams=num_beams, max_new_tokens=max_new_tokens,
                           eos_token_id=eos_token_id, pad_token_id=pad_token_id, length_penalty=length_penalty)
