"""Text bridge of the SCST step, restated (tokenizer objects are inputs).

TEST INFRASTRUCTURE (see oracle/__init__.py).

* split_and_decode_sections - reference modelling_longitudinal.py:413-457
* tokenize_prompt           - reference modelling_longitudinal.py:459-513
"""
from __future__ import annotations

import torch


def split_and_decode_sections(token_ids: torch.Tensor, special_token_ids, tokenizer):
    """Per row, section j = ids[prev_col : first column of special j] (a special
    found at column 0 or absent -> to the end; once prev_col has reached the end
    the remaining sections are '')."""
    _, seq_len = token_ids.shape
    sections = {k: [] for k in range(len(special_token_ids))}
    for row in token_ids:
        prev_col = 0
        for j, k in enumerate(special_token_ids):
            if prev_col >= seq_len:
                sections[j].append("")
                continue
            col = int((row == k).int().argmax())
            if col == 0:
                col = seq_len
            sections[j].append(tokenizer.decode(row[prev_col:col], skip_special_tokens=True))
            prev_col = col
    return tuple(sections.values())


def tokenize_prompt(previous_findings, previous_impression, tokenizer, max_len, add_bos_token_id=False):
    pf = ["[NPF]" if not i else i for i in previous_findings]
    pi = ["[NPI]" if not i else i for i in previous_impression]
    bos = tokenizer.bos_token if add_bos_token_id else ""
    texts = [f"[PMT]{i}[PMT-SEP]{j}{bos}" for i, j in zip(pf, pi)]
    out = tokenizer(texts, padding="longest", truncation=True, max_length=max_len, return_tensors="pt",
                    return_token_type_ids=False, add_special_tokens=False)
    ids, mask = out["input_ids"], out["attention_mask"]
    if ids.shape[1] == max_len:
        ids[:, -1] = torch.where(mask[:, -1] == 1, torch.full_like(ids[:, -1], tokenizer.bos_token_id), ids[:, -1])
    assert ids.shape[1] <= max_len
    return {"input_ids": ids, "attention_mask": mask}
