"""The SCST step around the rollout, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows reference modules/lightning_modules/longitudinal/scst/gen_prompt.py:
    scst_step       :174-259  (encode once -> sample -> reward -> greedy -> baseline -> advantage -> loss)
    sample          :261-329  (special_token_ids=[BOS,SEP], top-k 50, scores stacked [B,V,T])
    reinforce_loss  :331-366  (nll of sampled ids under log_softmax of the top-k-masked scores,
                               ignore_index=pad, sum over t, x advantage, mean over batch)
The LightningModule itself cannot be imported offline (lightning, torchmetrics,
pycocoevalcap, bert_score are absent), hence this restatement.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import cvt, decode, text

BOS, EOS, SEP, PAD, PMT_SEP = 1, 2, 3, 4, 9
SECTIONS = [0, 1, 0, 1]   # modelling_longitudinal.py:280-282


def reinforce_loss(logits: torch.Tensor, sampled_token_ids: torch.Tensor, reward: torch.Tensor, pad_token_id=PAD):
    """logits [B,V,T] (top-k-masked scores), sampled_token_ids [B,T], reward [B]."""
    loss = F.nll_loss(F.log_softmax(logits, dim=1), sampled_token_ids, ignore_index=pad_token_id, reduction="none")
    return (loss.sum(dim=-1) * reward).mean()


@dataclass
class ScstOut:
    loss: torch.Tensor
    reward: torch.Tensor        # advantage = sample reward - baseline (what the reference logs as 'reward')
    sample_reward: torch.Tensor
    baseline: torch.Tensor
    sample: decode.RolloutResult
    greedy: decode.RolloutResult
    sample_str: list
    baseline_str: list


def scst_step(sd, reward_fn, tokenizer, images, prompt_ids, label_texts, *, decoder_max_len=256, top_k=50,
              temperature=1.0, exp_noise=None, generator=None, depth=None, layers=6) -> ScstOut:
    """images [B,N,3,H,W]; prompt_ids [B,P] from tokenize_prompt(add_bos_token_id=True);
    label_texts: list[list[str]].  max_length = decoder_max_len + P counts HF's
    auto-prepended BOS, hence decoder_max_len - 1 new tokens."""
    kw = {} if depth is None else {"depth": depth}
    memory, memory_mask = cvt.encode_multi(sd, images, **kw)
    common = dict(mask_token_id=PAD, max_new_tokens=decoder_max_len - 1, eos_token_id=EOS, pad_token_id=PAD,
                  layers=layers, sections=SECTIONS)
    P = prompt_ids.shape[1]
    smp = decode.rollout(sd, memory, memory_mask, prompt_ids, special_token_ids=[BOS, SEP], do_sample=True,
                         top_k=top_k, temperature=temperature, exp_noise=exp_noise, generator=generator, **common)
    _, f, i = text.split_and_decode_sections(smp.sequences, [BOS, SEP, EOS], tokenizer)
    sample_str = [f"{a} {b}" for a, b in zip(f, i)]
    sample_reward = reward_fn(sample_str, label_texts)

    grd = decode.rollout(sd, memory, memory_mask, prompt_ids, special_token_ids=[PMT_SEP, BOS, SEP], do_sample=False,
                         **common)
    _, f, i = text.split_and_decode_sections(grd.sequences, [BOS, SEP, EOS], tokenizer)
    baseline_str = [f"{a} {b}" for a, b in zip(f, i)]
    baseline = reward_fn(baseline_str, label_texts)
    adv = sample_reward - baseline

    logits = torch.stack(smp.scores, dim=-1)
    loss = reinforce_loss(logits, smp.sequences[:, P:], adv)
    return ScstOut(loss, adv, sample_reward, baseline, smp, grd, sample_str, baseline_str)
