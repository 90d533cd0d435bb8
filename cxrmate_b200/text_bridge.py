"""The CPU text bridge of the SCST step: generated token ids -> report strings ->
reward-model token ids, i.e. `split_and_decode_sections` + `tokenizer.decode` of the decoder's byte-level BPE
(reference modelling_longitudinal.py:413-457, scst/gen_prompt.py:233-240,312-317) followed by CXR-BERT's
`batch_encode_plus` (tools/rewards/cxrbert.py:33-40,49-56).

It is the only serial CPU stage between the rollout and the reward.  Here it is batched: one host copy of the
sequences, column cuts found with vectorised torch ops, ONE `batch_decode` for all sections of all rows and ONE batched
WordPiece encode (the Rust tokenizers thread over the batch), results written into pinned buffers that go back to the
device asynchronously.  The label reports are tokenised on a worker thread while the GPU runs the rollout.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Sequence, Tuple

import numpy as np
import torch


class TextBridge:
    def __init__(self, decoder_tokenizer, reward_tokenizer, bos_token_id: int, sep_token_id: int, eos_token_id: int,
                 max_reward_len: int = 512):
        self.dec, self.rwd = decoder_tokenizer, reward_tokenizer
        self.specials = (bos_token_id, sep_token_id, eos_token_id)
        self.max_reward_len = max_reward_len
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._pinned = {}

    # ---- ids -> strings --------------------------------------------------------------------------------------------
    def split_ids(self, sequences: torch.Tensor) -> Tuple[List[List[int]], List[List[int]]]:
        """per row the id lists of the findings and the impression section: section j = ids[first col of special j-1 :
        first col of special j] with specials [BOS, SEP, EOS]; a special that is absent (or at column 0) cuts at the end,
        and once a cut has reached the end the remaining sections are empty (modelling_longitudinal.py:430-455)"""
        seq = sequences.detach().to("cpu")
        n, L = seq.shape
        cuts = []
        prev = torch.zeros(n, dtype=torch.int64)
        for tok in self.specials:
            col = (seq == tok).int().argmax(dim=1)
            col = torch.where(col == 0, torch.full_like(col, L), col)
            col = torch.where(prev >= L, prev, col)            # `continue` once the previous cut is at the end
            cuts.append((prev.clone(), col.clone()))
            prev = col
        rows = seq.tolist()
        f_lo, f_hi = cuts[1][0].tolist(), cuts[1][1].tolist()
        i_lo, i_hi = cuts[2][0].tolist(), cuts[2][1].tolist()
        findings = [rows[r][f_lo[r]:f_hi[r]] if f_lo[r] < L else [] for r in range(n)]
        impression = [rows[r][i_lo[r]:i_hi[r]] if i_lo[r] < L else [] for r in range(n)]
        return findings, impression

    def reports(self, sequences: torch.Tensor) -> List[str]:
        """f'{findings} {impression}' per row (scst/gen_prompt.py:237,317)"""
        f, i = self.split_ids(sequences)
        backend = getattr(self.dec, "backend_tokenizer", None)
        if backend is not None and not getattr(self.dec, "clean_up_tokenization_spaces", False):
            # the Rust tokenizer decodes the whole batch in one call (threads over the rows); `batch_decode` of the
            # transformers wrapper is a Python loop over `decode` that ends in the same routine
            txt = backend.decode_batch(f + i, skip_special_tokens=True)
        else:
            txt = self.dec.batch_decode(f + i, skip_special_tokens=True)
        n = len(f)
        return [f"{a} {b}" for a, b in zip(txt[:n], txt[n:])]

    # ---- strings -> reward-model ids -------------------------------------------------------------------------------
    def encode(self, texts: Sequence[str], key: str = "pred"):
        """(ids int32 [n, L] pinned, lens int32 [n] pinned) exactly as CXRBERTReward tokenises (padding='longest',
        truncation at max_position_embeddings)"""
        texts = list(texts)
        backend = getattr(self.rwd, "backend_tokenizer", None)
        pad_id = getattr(self.rwd, "pad_token_id", None)
        if backend is not None and pad_id is not None and len(texts) > 0:
            # straight to the Rust tokenizer: the transformers wrapper spends 4-5x the tokenisation time on building
            # BatchEncoding objects, Python-side padding and tensor conversion (same ids: tests/test_text_bridge.py)
            backend.enable_truncation(max_length=self.max_reward_len)
            backend.no_padding()
            encs = backend.encode_batch(texts, add_special_tokens=True)
            lens_np = np.fromiter((len(e.ids) for e in encs), dtype=np.int32, count=len(encs))
            n, L = len(encs), int(lens_np.max())
            ids_t, lens_t = self._buffers(key, n, L)
            ids_np = ids_t.numpy()
            ids_np.fill(pad_id)
            for r, e in enumerate(encs):
                ids_np[r, : lens_np[r]] = e.ids
            lens_t.numpy()[:] = lens_np
            return ids_t, lens_t
        enc = self.rwd(texts, add_special_tokens=True, padding="longest", return_tensors="pt", truncation=True,
                       max_length=self.max_reward_len)
        ids, lens = enc["input_ids"].to(torch.int32), enc["attention_mask"].sum(dim=1).to(torch.int32)
        out_ids, out_lens = self._buffers(key, ids.shape[0], ids.shape[1])
        out_ids.copy_(ids)
        out_lens.copy_(lens)
        return out_ids, out_lens

    def _buffers(self, key: str, n: int, L: int):
        """[n, L] int32 ids + [n] int32 lens, views of per-key pinned buffers when a GPU is present"""
        if not torch.cuda.is_available():
            return torch.empty(n, L, dtype=torch.int32), torch.empty(n, dtype=torch.int32)
        buf = self._pinned.get(key)
        if buf is None or buf[0].shape[0] < n or buf[0].shape[1] < L:
            buf = (torch.empty(max(n, 1), max(self.max_reward_len, L), dtype=torch.int32).pin_memory(),
                   torch.empty(max(n, 1), dtype=torch.int32).pin_memory())
            self._pinned[key] = buf
        return buf[0][:n, :L], buf[1][:n]

    def encode_async(self, texts: Sequence[str], key: str = "label"):
        """tokenise on the worker thread (labels, while the GPU is busy with the rollout); .result() -> (ids, lens)"""
        return self._pool.submit(self.encode, list(texts), key)

    def __call__(self, sequences: torch.Tensor):
        texts = self.reports(sequences)
        ids, lens = self.encode(texts)
        return texts, ids, lens
