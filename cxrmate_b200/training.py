"""The training half of the SCST step behind the reference's `reinforce_loss(...).backward()`
(reference modules/lightning_modules/longitudinal/scst/gen_prompt.py:331-366) and the teacher-forced step
(longitudinal/gt_prompt.py:186-249), with the decoder-gradient all-reduce of Lightning DDP.

The reference keeps the autograd graph of all 255 cached decode steps of the sampled rollout alive and backpropagates
through it.  Here the rollout is grad-free; its sampled sequence is run once, teacher-forced, through
`cxrm_train_step`, which returns the gradient of the same loss (eval-mode arithmetic is identical: a K/V-cached step
equals the corresponding column of a full causal pass).  Gradients land in ONE flat fp32 buffer; `all_reduce_grads`
averages it over the ranks with NCCL, stage by stage, overlapped with the backward pass.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .engine import Engine
from .modelling import position_ids_from_mask, token_ids_to_token_type_ids


def teacher_forced_inputs(sequences: torch.Tensor, prompt_len: int, special_token_ids, sections, pad_token_id: int):
    """sequences [R, P + T] (prompt + generated, PAD-filled after EOS) -> decoder inputs of the teacher-forced pass that
    scores every generated token: ids = seq[:, :-1]; targets = seq[:, 1:] at the generated positions, PAD elsewhere
    (ignored, as nll_loss(ignore_index=pad) does in the reference: scst/gen_prompt.py:356-362)."""
    ids = sequences[:, :-1]
    mask = (ids != pad_token_id).to(torch.int64)
    tt = token_ids_to_token_type_ids(ids, special_token_ids, sections)
    pos = position_ids_from_mask(mask)
    targets = sequences[:, 1:].clone()
    targets[:, : prompt_len - 1] = pad_token_id
    return ids, tt, pos, mask, targets


def pad_to_multiple(t: torch.Tensor, multiple: int, value: int) -> torch.Tensor:
    """right-pad the token axis so that rows x tokens divides by `multiple` (cxrm_train_step wants R * L % 8 == 0)"""
    R, L = t.shape
    need = 0
    while (R * (L + need)) % multiple:
        need += 1
    if need == 0:
        return t
    return torch.cat((t, torch.full((R, need), value, dtype=t.dtype, device=t.device)), dim=1)


def reinforce_backward(engine: Engine, sequences: torch.Tensor, prompt_len: int, advantage: torch.Tensor, *,
                       special_token_ids, sections, pad_token_id: int, top_k: int = 50, temperature: float = 1.0,
                       lora_only: bool = True, grads: Optional[torch.Tensor] = None, all_reduce: bool = False):
    """loss, grads = d/dtheta mean_b( -sum_t log p_topk(sampled[b, t]) * advantage[b] ) for the sample rows of a rollout.
    With all_reduce the flat gradient buffer is averaged over the ranks, bucket by bucket while the backward runs."""
    ids, tt, pos, mask, tgt = teacher_forced_inputs(sequences, prompt_len, special_token_ids, sections, pad_token_id)
    ids, tt, pos, mask = (pad_to_multiple(x, 8, v) for x, v in ((ids, pad_token_id), (tt, 0), (pos, 0), (mask, 0)))
    tgt = pad_to_multiple(tgt, 8, pad_token_id)
    kw = dict(loss_kind="reinforce", ignore_index=pad_token_id, advantage=advantage, top_k=top_k, temperature=temperature,
              lora_only=lora_only)
    return _staged(engine, (ids, tt, pos, mask, tgt), kw, grads, all_reduce)


def cross_entropy_backward(engine: Engine, decoder_input_ids, token_type_ids, position_ids, attention_mask, label_ids, *,
                           pad_token_id: int, lora_only: bool = False, grads: Optional[torch.Tensor] = None,
                           all_reduce: bool = False):
    """teacher-forced step: loss = cross_entropy(logits, label_ids, ignore_index=pad) (gt_prompt.py:231-236) and its gradients"""
    kw = dict(loss_kind="ce", ignore_index=pad_token_id, lora_only=lora_only)
    return _staged(engine, (decoder_input_ids, token_type_ids, position_ids, attention_mask, label_ids), kw, grads, all_reduce)


def _staged(engine, tensors, kw, grads, all_reduce):
    if not all_reduce or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return engine.train_step(*tensors, grads=grads, **kw)
    layout = engine.grad_layout(kw["lora_only"])
    world = dist.get_world_size()
    keep, loss, works = None, None, []
    for st in range(engine.train_stages):
        loss, grads = engine.train_step(*tensors, grads=grads, stage=st, _keep=keep, **kw)
        keep = engine._train_keep
        mine = [(o, ne) for _, o, ne, s_ in layout if s_ == st]
        if not mine:
            continue
        lo, hi = mine[0][0], mine[-1][0] + mine[-1][1]           # a stage owns one contiguous range
        bucket = grads[lo:hi]
        bucket.div_(world)                                         # DDP averages
        works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True))   # NCCL stream: overlaps the next stage
    for w in works:
        w.wait()
    return loss, grads
