"""BERT decoder (cross-attention, LoRA, tied LM head, KV cache) and the
CXR-BERT encoder, restated functionally from a state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  CPU, fp32, eval-mode semantics.

Follows HF transformers 5.5.0 `models/bert/modeling_bert.py`:
    BertEmbeddings.forward            :75-112   ((word + type) + position -> LayerNorm eps 1e-12)
    eager_attention_forward           :115-139  (QK^T * d^-0.5 + additive mask -> softmax -> V)
    BertSelfAttention / CrossAttention:143-284  (cache update; cross K/V computed once)
    BertSelfOutput / BertOutput       :287-298,344-356 (dense -> LayerNorm(x + residual))
    BertLayer.forward                 :379-421  (self -> cross -> FFN)
    BertLMPredictionHead              :471-494  (dense -> GELU -> LN -> tied decoder + bias)
and the reference's LoRA placement (modelling_longitudinal.py:163-170): rank 8,
alpha 32 on `attention.self.query|key` of every decoder layer (cross-attention
is NOT adapted because the regex is matched with re.fullmatch).
peft's forward is `base(x) + B(A(dropout(x))) * alpha / r`; dropout is the
identity in eval mode.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .weights import LORA_ALPHA, LORA_R

LN_EPS = 1e-12
HEADS = 12


def _lin(sd, p, x):
    y = F.linear(x, sd[p + ".weight"], sd[p + ".bias"])
    if p + ".lora_A.weight" in sd:
        y = y + F.linear(F.linear(x, sd[p + ".lora_A.weight"]), sd[p + ".lora_B.weight"]) * (LORA_ALPHA / LORA_R)
    return y


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], LN_EPS)


def _heads(x):
    b, t, h = x.shape
    return x.reshape(b, t, HEADS, h // HEADS).transpose(1, 2)


def _attend(q, k, v, add_mask):
    s = torch.matmul(q, k.transpose(2, 3)) * (q.shape[-1] ** -0.5)
    if add_mask is not None:
        s = s + add_mask
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2)
    return o.reshape(o.shape[0], o.shape[1], -1)


def embeddings(sd, prefix, ids, token_type_ids, position_ids):
    e = prefix + "embeddings."
    x = F.embedding(ids, sd[e + "word_embeddings.weight"])
    x = x + F.embedding(token_type_ids, sd[e + "token_type_embeddings.weight"])
    x = x + F.embedding(position_ids, sd[e + "position_embeddings.weight"])
    return _ln(sd, e + "LayerNorm", x)


@dataclass
class DecoderCache:
    """Self-attention K/V grow along the sequence; cross-attention K/V are
    written once (modeling_bert.py:247-262)."""
    self_k: list = field(default_factory=list)
    self_v: list = field(default_factory=list)
    cross_k: list = field(default_factory=list)
    cross_v: list = field(default_factory=list)

    def length(self) -> int:
        return 0 if not self.self_k else self.self_k[0].shape[2]


def decoder_hidden(sd, ids, token_type_ids, position_ids, key_mask, memory, memory_mask, cache: DecoderCache | None,
                   layers: int):
    """Run the decoder trunk on the tokens in `ids` [B,q].

    key_mask     [B, past+q] 1/0: validity of every self-attention key (the
                 `decoder_attention_mask` of the reference, full length).
    memory       [B,S,768] encoder states, memory_mask [B,S] bool or None.
    Returns last hidden states [B,q,768].
    """
    B, q = ids.shape
    past = cache.length() if cache is not None else 0
    neg = torch.finfo(torch.float32).min
    # causal AND key-padding (create_causal_mask + padding, modeling_bert.py:628-691)
    kpos = torch.arange(past + q, device=ids.device)
    qpos = torch.arange(past, past + q, device=ids.device)
    allowed = (kpos[None, :] <= qpos[:, None])[None, None] & key_mask.bool()[:, None, None, :]
    self_mask = torch.zeros(B, 1, q, past + q, device=ids.device).masked_fill(~allowed, neg)
    cross_mask = None
    if memory_mask is not None:
        cross_mask = torch.zeros(B, 1, 1, memory.shape[1], device=ids.device).masked_fill(~memory_mask.bool()[:, None, None, :], neg)

    x = embeddings(sd, "decoder.bert.", ids, token_type_ids, position_ids)
    for l in range(layers):
        p = f"decoder.bert.encoder.layer.{l}."
        qh = _heads(_lin(sd, p + "attention.self.query", x))
        kh = _heads(_lin(sd, p + "attention.self.key", x))
        vh = _heads(_lin(sd, p + "attention.self.value", x))
        if cache is not None:
            if len(cache.self_k) <= l:
                cache.self_k.append(kh)
                cache.self_v.append(vh)
            else:
                cache.self_k[l] = torch.cat((cache.self_k[l], kh), dim=2)
                cache.self_v[l] = torch.cat((cache.self_v[l], vh), dim=2)
            kh, vh = cache.self_k[l], cache.self_v[l]
        a = _attend(qh, kh, vh, self_mask)
        x = _ln(sd, p + "attention.output.LayerNorm", _lin(sd, p + "attention.output.dense", a) + x)

        qh = _heads(_lin(sd, p + "crossattention.self.query", x))
        if cache is not None and len(cache.cross_k) > l:
            kh, vh = cache.cross_k[l], cache.cross_v[l]
        else:
            kh = _heads(_lin(sd, p + "crossattention.self.key", memory))
            vh = _heads(_lin(sd, p + "crossattention.self.value", memory))
            if cache is not None:
                cache.cross_k.append(kh)
                cache.cross_v.append(vh)
        a = _attend(qh, kh, vh, cross_mask)
        x = _ln(sd, p + "crossattention.output.LayerNorm", _lin(sd, p + "crossattention.output.dense", a) + x)

        h = F.gelu(_lin(sd, p + "intermediate.dense", x))
        x = _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", h) + x)
    return x


def lm_head(sd, x):
    t = "decoder.cls.predictions.transform."
    y = _ln(sd, t + "LayerNorm", F.gelu(_lin(sd, t + "dense", x)))
    return F.linear(y, sd["decoder.bert.embeddings.word_embeddings.weight"], sd["decoder.cls.predictions.bias"])


def decoder_logits(sd, ids, token_type_ids, position_ids, key_mask, memory, memory_mask, cache=None, layers=6,
                   last_only=False):
    x = decoder_hidden(sd, ids, token_type_ids, position_ids, key_mask, memory, memory_mask, cache, layers)
    if last_only:
        x = x[:, -1:]
    return lm_head(sd, x)


# --------------------------------------------------------------------------
# CXR-BERT (reward model): BERT-base encoder + CLS projection head.
# Call sites: reference tools/rewards/cxrbert.py:42-63 (input_ids + attention_mask,
# element [2] of the tuple = projected CLS embedding).
# --------------------------------------------------------------------------

def cxrbert_hidden(sd, ids, attention_mask, layers=12):
    B, T = ids.shape
    neg = torch.finfo(torch.float32).min
    add_mask = torch.zeros(B, 1, 1, T, device=ids.device).masked_fill(~attention_mask.bool()[:, None, None, :], neg)
    pos = torch.arange(T, device=ids.device)[None].expand(B, T)
    x = embeddings(sd, "bert.", ids, torch.zeros_like(ids), pos)
    for l in range(layers):
        p = f"bert.encoder.layer.{l}."
        qh = _heads(_lin(sd, p + "attention.self.query", x))
        kh = _heads(_lin(sd, p + "attention.self.key", x))
        vh = _heads(_lin(sd, p + "attention.self.value", x))
        a = _attend(qh, kh, vh, add_mask)
        x = _ln(sd, p + "attention.output.LayerNorm", _lin(sd, p + "attention.output.dense", a) + x)
        h = F.gelu(_lin(sd, p + "intermediate.dense", x))
        x = _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", h) + x)
    return x


def cxrbert_cls_projection(sd, ids, attention_mask, layers=12):
    x = cxrbert_hidden(sd, ids, attention_mask, layers)[:, 0]
    h = "cls_projection_head."
    y = F.gelu(_lin(sd, h + "dense_to_hidden", x))
    y = _ln(sd, h + "LayerNorm", y)
    return _lin(sd, h + "dense_to_output", y)
