"""CXRBERTReward restated (reference tools/rewards/cxrbert.py:23-73).

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference tokenises predictions and the flattened labels with the CXR-BERT
tokenizer (`padding='longest'`, truncation to `max_position_embeddings` = 512),
runs the hub model twice with `output_cls_projected_embedding=True` and returns
`cosine_similarity(pred[2], label[2])`.  The hub code is not reachable offline,
so the trunk is `oracle.bert.cxrbert_cls_projection` (parity unpinned against
the hub implementation; the call contract above is what is reproduced).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import bert


def reward_from_ids(sd, pred_ids, pred_mask, label_ids, label_mask, layers=12) -> torch.Tensor:
    """ids/masks are the `batch_encode_plus(..., padding='longest')` outputs."""
    pe = bert.cxrbert_cls_projection(sd, pred_ids, pred_mask, layers)
    le = bert.cxrbert_cls_projection(sd, label_ids, label_mask, layers)
    return F.cosine_similarity(pe, le)


class CXRBERTReward:
    """Same call surface as the reference class: `reward(predictions: list[str],
    labels: list[list[str]]) -> FloatTensor[B]` (asserts copied in meaning from
    cxrbert.py:24-28)."""

    def __init__(self, sd, tokenizer, layers=12, max_len=512):
        self.sd, self.tokenizer, self.layers, self.max_len = sd, tokenizer, layers, max_len

    def __call__(self, predictions, labels):
        return self.reward(predictions, labels)

    def _tok(self, texts):
        out = self.tokenizer(texts, add_special_tokens=True, padding="longest", return_tensors="pt",
                             truncation=True, max_length=self.max_len)
        return out["input_ids"], out["attention_mask"]

    def reward(self, predictions, labels):
        assert isinstance(predictions, list) and all(isinstance(i, str) for i in predictions)
        assert isinstance(labels, list) and all(isinstance(i, list) for i in labels)
        assert all(isinstance(j, str) for i in labels for j in i)
        with torch.no_grad():
            pi, pm = self._tok(predictions)
            li, lm = self._tok([j for i in labels for j in i])
            return reward_from_ids(self.sd, pi, pm, li, lm, self.layers)
