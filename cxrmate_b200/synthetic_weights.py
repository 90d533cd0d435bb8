"""Deterministic synthetic weights in the reference's state_dict naming.

An INPUT generator (no model arithmetic): used by bench.py and smoke() for the
engine, and re-exported as `oracle.weights` for the CPU oracle, so that both
sides of every parity test see the same tensors.

Names and shapes follow the instantiated reference model
(modules/transformers/longitudinal_model/modelling_longitudinal.py:93-171 ->
SURVEY.md Appendix D).  No checkpoint exists offline, so every tensor is drawn
from a seeded CPU generator; every bias / LayerNorm / BatchNorm tensor
(including the running statistics) and the LoRA B matrices are perturbed away
from their "identity" initial values, because zero biases, unit norms and a
zero LoRA-B would hide indexing bugs.

The draw order is the insertion order of the dict built below; it must never
change, because tests/golden/*.npz were produced from it.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

CVT_EMBED_DIM = (64, 192, 384)
CVT_DEPTH = (1, 4, 16)
CVT_HEADS = (1, 3, 6)
CVT_PATCH = (7, 3, 3)
CVT_STRIDE = (4, 2, 2)
CVT_PAD = (2, 1, 1)
DEC_LAYERS = 6
DEC_HIDDEN = 768
DEC_FFN = 3072
DEC_VOCAB = 30000
DEC_MAXPOS = 512
LORA_R = 8
LORA_ALPHA = 32
RWD_LAYERS = 12
RWD_VOCAB = 30522
RWD_PROJ = 128


class _Draw:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)
        self.sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def normal(self, name, shape, std, mean=0.0):
        self.sd[name] = torch.randn(shape, generator=self.g) * std + mean

    def uniform(self, name, shape, lo, hi):
        self.sd[name] = torch.rand(shape, generator=self.g) * (hi - lo) + lo

    def linear(self, prefix, n_out, n_in, std=0.02, bias=True):
        self.normal(prefix + ".weight", (n_out, n_in), std)
        if bias:
            self.normal(prefix + ".bias", (n_out,), 0.02)

    def layernorm(self, prefix, n):
        self.normal(prefix + ".weight", (n,), 0.05, mean=1.0)
        self.normal(prefix + ".bias", (n,), 0.05)


def make_encoder_weights(d: _Draw, depth=CVT_DEPTH) -> None:
    c_in = 3
    for s, C in enumerate(CVT_EMBED_DIM):
        p = f"encoder.cvt.encoder.stages.{s}."
        k = CVT_PATCH[s]
        # fan-in scaled so that activations entering the LayerNorm are O(1)
        d.normal(p + "embedding.convolution_embeddings.projection.weight", (C, c_in, k, k), (c_in * k * k) ** -0.5)
        d.normal(p + "embedding.convolution_embeddings.projection.bias", (C,), 0.02)
        d.layernorm(p + "embedding.convolution_embeddings.normalization", C)
        if s == 2:
            d.normal(p + "cls_token", (1, 1, C), 1.0)
        for i in range(depth[s]):
            q = p + f"layers.{i}."
            for nm in ("query", "key", "value"):
                cp = q + f"attention.attention.convolution_projection_{nm}.convolution_projection."
                d.normal(cp + "convolution.weight", (C, 1, 3, 3), 1.0 / 3.0)
                d.normal(cp + "normalization.weight", (C,), 0.05, mean=1.0)
                d.normal(cp + "normalization.bias", (C,), 0.05)
                d.normal(cp + "normalization.running_mean", (C,), 0.1)
                d.uniform(cp + "normalization.running_var", (C,), 0.5, 1.5)
                d.sd[cp + "normalization.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
            for nm in ("query", "key", "value"):
                # larger than HF's 0.02 so that the softmax is not uniform
                d.linear(q + f"attention.attention.projection_{nm}", C, C, std=2.0 * C ** -0.5)
            d.linear(q + "attention.output.dense", C, C, std=0.5 * C ** -0.5)
            d.linear(q + "intermediate.dense", 4 * C, C, std=C ** -0.5)
            d.linear(q + "output.dense", C, 4 * C, std=0.5 * (4 * C) ** -0.5)
            d.layernorm(q + "layernorm_before", C)
            d.layernorm(q + "layernorm_after", C)
        c_in = C
    d.layernorm("encoder.projection_head.layer_norm", CVT_EMBED_DIM[-1])
    d.normal("encoder.projection_head.projection.weight", (DEC_HIDDEN, CVT_EMBED_DIM[-1]), CVT_EMBED_DIM[-1] ** -0.5)


def _bert_layer(d: _Draw, p: str, H: int, F: int, cross: bool, lora: bool) -> None:
    for nm in ("query", "key", "value"):
        d.linear(p + f"attention.self.{nm}", H, H, std=2.0 * H ** -0.5)
        if lora and nm in ("query", "key"):
            d.normal(p + f"attention.self.{nm}.lora_A.weight", (LORA_R, H), H ** -0.5)
            d.normal(p + f"attention.self.{nm}.lora_B.weight", (H, LORA_R), 0.05)
    d.linear(p + "attention.output.dense", H, H, std=0.5 * H ** -0.5)
    d.layernorm(p + "attention.output.LayerNorm", H)
    if cross:
        for nm in ("query", "key", "value"):
            d.linear(p + f"crossattention.self.{nm}", H, H, std=2.0 * H ** -0.5)
        d.linear(p + "crossattention.output.dense", H, H, std=0.5 * H ** -0.5)
        d.layernorm(p + "crossattention.output.LayerNorm", H)
    d.linear(p + "intermediate.dense", F, H, std=H ** -0.5)
    d.linear(p + "output.dense", H, F, std=0.5 * F ** -0.5)
    d.layernorm(p + "output.LayerNorm", H)


def make_decoder_weights(d: _Draw, lora: bool = True, layers: int = DEC_LAYERS, vocab: int = DEC_VOCAB) -> None:
    H = DEC_HIDDEN
    e = "decoder.bert.embeddings."
    # std 0.1: the tied LM head then yields logits with std ~2.8, i.e. clear
    # top-1/top-2 margins for greedy parity (SURVEY.md section 7 "hard parts")
    d.normal(e + "word_embeddings.weight", (vocab, H), 0.1)
    d.normal(e + "position_embeddings.weight", (DEC_MAXPOS, H), 0.1)
    d.normal(e + "token_type_embeddings.weight", (2, H), 0.1)
    d.layernorm(e + "LayerNorm", H)
    for l in range(layers):
        _bert_layer(d, f"decoder.bert.encoder.layer.{l}.", H, DEC_FFN, cross=True, lora=lora)
    d.linear("decoder.cls.predictions.transform.dense", H, H, std=H ** -0.5)
    d.layernorm("decoder.cls.predictions.transform.LayerNorm", H)
    d.normal("decoder.cls.predictions.bias", (vocab,), 0.5)
    # tied tensors (finding 3): same storage, listed for completeness
    d.sd["decoder.cls.predictions.decoder.weight"] = d.sd[e + "word_embeddings.weight"]
    d.sd["decoder.cls.predictions.decoder.bias"] = d.sd["decoder.cls.predictions.bias"]


def make_cxrmate_weights(seed: int = 0, lora: bool = True, depth=CVT_DEPTH, layers: int = DEC_LAYERS,
                         vocab: int = DEC_VOCAB) -> "OrderedDict[str, torch.Tensor]":
    """Full encoder-decoder state dict (fp32, CPU).

    `depth`, `layers` and `vocab` exist so that CPU-side unit tests can build a
    shallow model quickly; the named architecture is the default.
    """
    d = _Draw(seed)
    make_encoder_weights(d, depth)
    make_decoder_weights(d, lora, layers, vocab)
    return d.sd


def make_cxrbert_weights(seed: int = 1, layers: int = RWD_LAYERS, vocab: int = RWD_VOCAB) -> "OrderedDict[str, torch.Tensor]":
    """CXR-BERT-specialized architecture: BERT-base trunk (no pooler) + CLS
    projection head 768 -> 128 -> GELU -> LN -> 128 (call sites: reference
    tools/rewards/cxrbert.py:42-63)."""
    d = _Draw(seed)
    H = DEC_HIDDEN
    e = "bert.embeddings."
    d.normal(e + "word_embeddings.weight", (vocab, H), 0.1)
    d.normal(e + "position_embeddings.weight", (DEC_MAXPOS, H), 0.1)
    d.normal(e + "token_type_embeddings.weight", (2, H), 0.1)
    d.layernorm(e + "LayerNorm", H)
    for l in range(layers):
        _bert_layer(d, f"bert.encoder.layer.{l}.", H, DEC_FFN, cross=False, lora=False)
    d.linear("cls_projection_head.dense_to_hidden", RWD_PROJ, H, std=H ** -0.5)
    d.layernorm("cls_projection_head.LayerNorm", RWD_PROJ)
    d.linear("cls_projection_head.dense_to_output", RWD_PROJ, RWD_PROJ, std=RWD_PROJ ** -0.5)
    return d.sd


def count_params(sd) -> int:
    seen, n = set(), 0
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            continue
        if v.data_ptr() in seen:
            continue
        seen.add(v.data_ptr())
        n += v.numel()
    return n
