"""Drop-in for the reference's reward callable (reference tools/rewards/cxrbert.py:9-73).

    reward = CXRBERTReward(device, engine=eng, tokenizer=cxr_bert_tokenizer)
    r = reward(predictions: list[str], labels: list[list[str]])      # FloatTensor[B] on `device`

Same call surface, asserts and tokenisation call as the reference class; the model forward
(`self.model(..., output_cls_projected_embedding=True)[2]`) and the cosine are executed by libcxrm.so
(`cxrm_reward_embed`, `cxrm_cosine`).  The reference builds tokenizer and model from the hub inside `__init__`; there
is no network here, so both are passed in (the engine holds the CXR-BERT weights under the `reward.` prefix).
"""
from __future__ import annotations

import torch

from .engine import Engine


class CXRBERTReward:

    def __init__(self, device, engine: Engine = None, tokenizer=None, max_position_embeddings: int = 512):
        if engine is None or tokenizer is None:
            raise ValueError("CXRBERTReward needs the engine holding the reward weights and the CXR-BERT tokenizer "
                             "(the hub checkpoint microsoft/BiomedVLP-CXR-BERT-specialized is not reachable offline)")
        self.device = torch.device(device)
        self.engine = engine
        self.tokenizer = tokenizer
        self.max_position_embeddings = min(max_position_embeddings, engine.cfg.rwd_max_len)

    def __call__(self, predictions, labels):
        return self.reward(predictions, labels)

    def _embed(self, texts):
        tok = self.tokenizer(texts, add_special_tokens=True, padding="longest", return_tensors="pt", truncation=True,
                             max_length=self.max_position_embeddings)      # cxrbert.py:33-40,49-56
        ids = tok["input_ids"].to(self.device, non_blocking=True)
        lens = tok["attention_mask"].sum(dim=1).to(self.device, non_blocking=True)
        out = []
        cap = max(1, (self.engine.cfg.rwd_max_seqs * self.engine.cfg.rwd_max_len) // ids.shape[1])
        for i in range(0, ids.shape[0], cap):                               # the engine's reward batch is bounded
            out.append(self.engine.reward_embed(ids[i:i + cap], lens[i:i + cap]))
        return torch.cat(out)

    def reward(self, predictions, labels):
        assert isinstance(predictions, list), '"predictions" must be a list of strings.'
        assert all(isinstance(i, str) for i in predictions), 'Each element of "predictions" must be a string.'
        assert isinstance(labels, list), '"labels" must be a list of lists, where each sub-list has a multiple strings.'
        assert all(isinstance(i, list) for i in labels), 'Each element of "labels" must be a list of strings.'
        assert all(isinstance(j, str) for i in labels for j in i), 'each sub-list must have one or more strings.'
        with torch.no_grad():
            pe = self._embed(predictions)
            le = self._embed([j for i in labels for j in i])
            if pe.shape[0] != le.shape[0]:     # torch.nn.functional.cosine_similarity would raise on the same input
                raise RuntimeError(f"The size of tensor a ({pe.shape[0]}) must match the size of tensor b ({le.shape[0]})")
            return self.engine.cosine(pe, le)
