"""Reference-facing Python surface of the engine.

`CXRMateEngineModel` mirrors the call signatures of the reference's
`LongitudinalPromptMultiCXREncoderDecoderModel` (reference
modules/transformers/longitudinal_model/modelling_longitudinal.py:93-513), of
`MultiCXREncoderDecoderModel` (multi_model/modelling_multi.py:90-261) and of
`SingleCXREncoderDecoderModel` (single_model/modelling_single.py:81-249) for the
rollout path: `.encoder(pixel_values)`, `.generate(...)`, `.forward(...)` and the
tokenisation helpers.  Every tensor operation is executed by libcxrm.so; this
file only marshals arguments (and raises the reference's errors).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from .engine import Engine, RolloutOutput


TOPK_CAP = 64   # survivors recorded per (row, step): include/cxrm.h cxrm_rollout_args.topk_idx


class DecoderPast:
    """`past_key_values` of CXRMateEngineModel.forward: ids / token types / positions of the cached tokens."""

    def __init__(self, ids, tt, pos):
        self.ids, self.tt, self.pos = ids, tt, pos

    @property
    def length(self) -> int:
        return int(self.ids.shape[1])

    def get_seq_length(self, layer_idx: int = 0) -> int:     # transformers.Cache surface used by callers
        return self.length

    def __len__(self):
        return self.length

    def __bool__(self):
        return self.length > 0


class ModelOutput(dict):
    """dict with attribute access, like transformers.modeling_outputs.ModelOutput."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)


def token_ids_to_token_type_ids(token_ids: torch.Tensor, special_token_ids, token_type_id_sections=None):
    """Vectorised form of modelling_longitudinal.py:297-338 (columns strictly after the first occurrence of the
    i-th special token get sections[i+1]; a token that is absent, at column 0, or last has no effect)."""
    sections = token_type_id_sections if token_type_id_sections is not None else list(range(len(special_token_ids) + 1))
    B, L = token_ids.shape
    tt = torch.full_like(token_ids, sections[0], dtype=torch.long)
    col = torch.arange(L, device=token_ids.device)[None]
    for i, tok in enumerate(special_token_ids):
        first = (token_ids == tok).int().argmax(dim=1) + 1
        ok = (first != 1) & (first < L)
        tt = torch.where(ok[:, None] & (col >= first[:, None]), torch.full_like(tt, sections[i + 1]), tt)
    return tt


def token_ids_to_token_type_ids_past(token_ids: torch.Tensor, special_token_ids, token_type_id_sections=None):
    """modelling_longitudinal.py:340-364."""
    sections = token_type_id_sections if token_type_id_sections is not None else list(range(len(special_token_ids) + 1))
    tt = torch.full((token_ids.shape[0], 1), sections[0], dtype=torch.long, device=token_ids.device)
    prev = token_ids[:, :-1]
    for i, tok in enumerate(special_token_ids):
        tt[torch.any(prev == tok, dim=1, keepdim=True)] = sections[i + 1]
    return tt


def position_ids_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """modelling_longitudinal.py:275-277."""
    return torch.nn.functional.relu(torch.cumsum(mask.to(torch.int64), dim=1) - 1)


class _Encoder:
    """`model.encoder(pixel_values)` -> ModelOutput(last_hidden_state, attention_mask)."""

    def __init__(self, model: "CXRMateEngineModel"):
        self.model = model

    def __call__(self, pixel_values: torch.Tensor, output_hidden_states=None, return_dict=None, **kwargs):
        m = self.model
        single = pixel_values.dim() == 4
        px = pixel_values[:, None] if single else pixel_values
        memory, mask = m.engine.encode(px.to(device=m.device, dtype=torch.float32))
        m._encoded_token = object()
        memory._cxrm_token = m._encoded_token      # lets generate() skip re-uploading the engine's own result
        if single or m.variant == "single":
            return ModelOutput(last_hidden_state=memory)          # modelling_single.py:74-78: no mask
        return ModelOutput(last_hidden_state=memory, attention_mask=mask)


class CXRMateEngineModel:
    """Drop-in for the rollout path of the three reference encoder-decoder classes.

    variant: 'longitudinal' (prompt, positions from the mask, sections [0,1,0,1]),
             'multi' / 'single' (prompt is [BOS]; default sections; no mask_token_id).
    """

    main_input_name = "pixel_values"

    def __init__(self, engine: Engine, variant: str = "longitudinal", bos_token_id: int = 1, eos_token_id: int = 2,
                 pad_token_id: int = 4):
        assert variant in ("longitudinal", "multi", "single")
        self.engine = engine
        self.variant = variant
        self.device = engine.device
        self.config = SimpleNamespace(bos_token_id=bos_token_id, eos_token_id=eos_token_id, pad_token_id=pad_token_id,
                                      decoder=SimpleNamespace(vocab_size=engine.cfg.vocab))
        self.encoder = _Encoder(self)
        self._encoded_token = None
        self._kv_token = None

    @classmethod
    def from_state_dict(cls, state_dict, reward_state_dict=None, variant="longitudinal", **engine_kwargs):
        if reward_state_dict is None:
            engine_kwargs.setdefault("rwd_layers", 0)
        eng = Engine(**engine_kwargs)
        eng.load_state_dict(state_dict)
        if reward_state_dict is not None:
            eng.load_state_dict(reward_state_dict, prefix="reward.")
        eng.finalize()
        return cls(eng, variant)

    def eval(self):
        return self

    # ---- helpers with the reference's names --------------------------------------------------------
    token_ids_to_token_type_ids = staticmethod(token_ids_to_token_type_ids)
    token_ids_to_token_type_ids_past = staticmethod(token_ids_to_token_type_ids_past)

    def tokenize_prompt(self, previous_findings, previous_impression, tokenizer, max_len, add_bos_token_id=False):
        """modelling_longitudinal.py:459-513."""
        pf = ["[NPF]" if not i else i for i in previous_findings]
        pi = ["[NPI]" if not i else i for i in previous_impression]
        bos = tokenizer.bos_token if add_bos_token_id else ""
        texts = [f"[PMT]{i}[PMT-SEP]{j}{bos}" for i, j in zip(pf, pi)]
        out = tokenizer(texts, padding="longest", truncation=True, max_length=max_len, return_tensors="pt",
                        return_token_type_ids=False, add_special_tokens=False).to(self.device)
        if out.input_ids.shape[1] == max_len:
            out.input_ids[:, -1] = torch.where(out.attention_mask[:, -1] == 1, tokenizer.bos_token_id,
                                               out.input_ids[:, -1])
        assert out.input_ids.shape[1] <= max_len
        return {"input_ids": out.input_ids, "attention_mask": out.attention_mask}

    def tokenize_report_teacher_forcing(self, findings, impression, tokenizer, max_len):
        """modelling_longitudinal.py:366-411."""
        report = [f"{tokenizer.bos_token}{i}{tokenizer.sep_token}{j}{tokenizer.eos_token}" for i, j in
                  zip(findings, impression)]
        tok = tokenizer(report, padding="longest", truncation=True, max_length=max_len + 1, return_tensors="pt",
                        return_token_type_ids=False, add_special_tokens=False).to(self.device)
        return {"label_ids": tok["input_ids"][:, 1:].detach().clone(), "decoder_input_ids": tok["input_ids"][:, :-1],
                "decoder_attention_mask": tok["attention_mask"][:, 1:]}

    def split_and_decode_sections(self, token_ids, special_token_ids, tokenizer):
        """modelling_longitudinal.py:413-457, with one host copy instead of one `.item()` per row and section."""
        ids = token_ids.detach().cpu()
        n_rows, seq_len = ids.shape
        sections = {k: [] for k in range(len(special_token_ids))}
        first = {k: (ids == tok).int().argmax(dim=1).tolist() for k, tok in enumerate(special_token_ids)}
        rows = ids.tolist()
        for r in range(n_rows):
            prev = 0
            for j in range(len(special_token_ids)):
                if prev >= seq_len:
                    sections[j].append("")
                    continue
                col = first[j][r] or seq_len
                sections[j].append(tokenizer.decode(rows[r][prev:col], skip_special_tokens=True))
                prev = col
        return tuple(sections.values())

    # ---- generation --------------------------------------------------------------------------------
    def _ensure_cross_kv(self, pixel_values, encoder_outputs):
        eng = self.engine
        if encoder_outputs is None:
            if pixel_values is None:
                raise ValueError("You have to specify pixel_values")
            encoder_outputs = self.encoder(pixel_values)
        memory = encoder_outputs["last_hidden_state"] if isinstance(encoder_outputs, dict) else encoder_outputs[0]
        mask = encoder_outputs.get("attention_mask") if isinstance(encoder_outputs, dict) else None
        token = getattr(memory, "_cxrm_token", None)
        if token is not None and token is self._encoded_token:
            if self._kv_token is not token:
                eng.prefill_cross_kv()                 # the engine still holds this encode() result
                self._kv_token = token
        else:
            eng.prefill_cross_kv(memory.to(self.device), None if mask is None else mask.to(self.device))
            self._kv_token = None
        return memory.shape[0]

    def _generate(self, pixel_values=None, encoder_outputs=None, decoder_input_ids=None, input_ids=None,
                 special_token_ids=None, mask_token_id=None, max_length=None, max_new_tokens=None, bos_token_id=None,
                 eos_token_id=None, pad_token_id=None, num_beams=1, do_sample=False, top_k=50, top_p=1.0,
                 temperature=1.0, output_scores=False, return_dict_in_generate=False, use_cache=True,
                 exp_noise: Optional[torch.Tensor] = None, seed: int = 0, **kwargs):
        """Greedy or top-k multinomial rollout with the keyword surface the reference passes to HF `generate`
        (scst/gen_prompt.py:206-224,279-300; single.py:483-493; multi.py:218-228).

        Returns `sequences` [B, 1+P+t] when a prompt was given (HF prepends decoder_start_token_id = BOS to
        user-supplied decoder ids that do not start with it; the reference strips it again) and [B, 1+t] for the
        prompt-free variants; with return_dict_in_generate a mapping that also holds `scores` (tuple of t fp32
        [B,V] tensors, top-k-masked for sampling) when output_scores is set.
        """
        if num_beams != 1 and do_sample:
            raise NotImplementedError("beam-sample is not used by the reference (test_step: num_beams > 1, do_sample False)")
        if do_sample and top_k is not None and top_k > TOPK_CAP and output_scores:
            raise ValueError(f"output_scores with top_k > {TOPK_CAP}: the engine records at most {TOPK_CAP} survivors per step")
        if top_p != 1.0:
            raise NotImplementedError("top_p != 1.0 is not used by the reference's SCST recipe")
        cfg = self.config
        bos = cfg.bos_token_id if bos_token_id is None else bos_token_id
        eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
        pad = cfg.pad_token_id if pad_token_id is None else pad_token_id
        B = self._ensure_cross_kv(pixel_values, encoder_outputs)
        prompt = decoder_input_ids if decoder_input_ids is not None else input_ids
        had_prompt = prompt is not None
        if prompt is None:
            prompt = torch.full((B, 1), bos, dtype=torch.int64, device=self.device)
        prompt = prompt.to(self.device)
        # prepare_inputs_for_generation strips one leading BOS column when every row has it (:270-271)
        auto_bos = not bool(torch.all(prompt[:, 0] == bos))
        P = prompt.shape[1]
        total_in = P + (1 if auto_bos else 0)          # what HF counts against max_length
        if max_new_tokens is None:
            if max_length is None:
                raise ValueError("max_length or max_new_tokens is required")
            max_new_tokens = max_length - total_in
        if max_new_tokens < 1:
            raise ValueError("max_length leaves no room for new tokens")
        if self.variant == "longitudinal":
            sections = [0, 1, 0, 1][: len(special_token_ids) + 1]   # modelling_longitudinal.py:280-282
            if mask_token_id is None:
                raise TypeError("generate() missing mask_token_id (required by the longitudinal model)")
        else:
            sections = list(range(len(special_token_ids) + 1))      # modelling_multi.py:248-250
            mask_token_id = None
        if not auto_bos and self.variant == "longitudinal":
            prompt_dec = prompt[:, 1:] if P > 1 else prompt
        else:
            prompt_dec = prompt
        if self.variant != "longitudinal":
            prompt_dec = prompt                                       # [BOS] is the real first decoder token
        if num_beams != 1:
            # test_step of the reference (gt_prompt.py:344-362, single.py:552-562): HF beam search, default flags
            bo = self.engine.rollout_beam(prompt_dec, num_beams=num_beams, max_new_tokens=max_new_tokens, eos_token_id=eos,
                                          pad_token_id=pad, mask_token_id=mask_token_id, special=special_token_ids,
                                          sections=sections, length_penalty=kwargs.get("length_penalty", 1.0))
            seq = bo.sequences
            if auto_bos or (self.variant == "longitudinal" and prompt_dec.shape[1] != P):
                seq = torch.cat((torch.full((B, 1), bos, dtype=seq.dtype, device=seq.device), seq), dim=1)
            if not return_dict_in_generate:
                return seq
            return ModelOutput(sequences=seq, sequences_scores=bo.scores, steps=bo.steps)
        mode = "sample" if do_sample else "greedy"
        kw = dict(special_sample=special_token_ids, sections_sample=sections) if do_sample else \
            dict(special_greedy=special_token_ids, sections_greedy=sections)
        out: RolloutOutput = self.engine.rollout(
            prompt_dec, mode=mode, max_new_tokens=max_new_tokens, eos_token_id=eos, pad_token_id=pad,
            mask_token_id=mask_token_id, top_k=top_k if do_sample else 0, temperature=temperature,
            exp_noise=exp_noise, seed=seed, **kw)
        seq = out.sequences[:, : prompt_dec.shape[1] + out.steps]
        if auto_bos or (self.variant == "longitudinal" and prompt_dec.shape[1] != P):
            seq = torch.cat((torch.full((B, 1), bos, dtype=seq.dtype, device=seq.device), seq), dim=1)
        if not return_dict_in_generate:
            return seq
        res = ModelOutput(sequences=seq)
        if output_scores:
            res["scores"] = self._dense_scores(out, do_sample)
        res["logprobs"] = out.logprobs[:, : out.steps]
        res["steps"] = out.steps
        res["rollout"] = out
        return res

    # The reference's SCST sampler calls `generate.__wrapped__(model, ...)` to step around HF's @torch.no_grad()
    # (scst/gen_prompt.py:279, scst/gt_prompt.py:162).  The engine's rollout is grad-free by construction (the log-probs
    # the REINFORCE loss needs are recorded by the sampling head), so the wrapped and the unwrapped call are the same
    # function; `generate.__wrapped__` exists so that the reference's call site works unchanged.
    generate = torch.no_grad()(_generate)

    def _dense_scores(self, out: RolloutOutput, sampled: bool):
        """HF's `scores`: one [B,V] fp32 tensor per executed step, -inf outside the top-k survivors.  One scatter for
        all steps (a [steps, B, V] tensor), returned as a tuple of its per-step views.  Raises when a step had more
        survivors than the engine records (ties at the k-th value beyond TOPK_CAP slots): the dense scores would
        silently miss entries otherwise."""
        if not sampled:
            raise NotImplementedError("dense greedy scores are not kept (only the last step's logits are)")
        V = self.engine.cfg.vocab
        T = out.steps
        cap = out.topk_idx.shape[-1]
        cnt = out.topk_cnt[:, :T]
        if T and int(cnt.max()) > cap:
            r, t = divmod(int(cnt.argmax()), T)
            raise RuntimeError(f"step {t} of row {r} kept {int(cnt.max())} top-k survivors (ties at the k-th value); only "
                               f"{cap} are recorded, so `scores` cannot be rebuilt exactly - use `logprobs`")
        Bn = out.topk_idx.shape[0]
        dense = torch.full((T, Bn, V), float("-inf"), device=out.topk_idx.device)
        valid = torch.arange(cap, device=dense.device)[None, None] < cnt.t()[:, :, None]          # [T, B, cap]
        idx = out.topk_idx[:, :T].permute(1, 0, 2).long()
        val = out.topk_val[:, :T].permute(1, 0, 2)
        tb = torch.nonzero(valid, as_tuple=True)
        dense[tb[0], tb[1], idx[valid]] = val[valid]
        return tuple(dense.unbind(0))

    # ---- teacher-forced forward --------------------------------------------------------------------
    def forward(self, pixel_values=None, decoder_input_ids=None, decoder_attention_mask=None, encoder_outputs=None,
                decoder_token_type_ids=None, decoder_position_ids=None, past_key_values=None, use_cache=None,
                labels=None, return_dict=True, **kwargs):
        """Decoder forward (modelling_longitudinal.py:173-249; multi: modelling_multi.py:150-227; single:
        modelling_single.py:138-215) -> ModelOutput(logits [B,q,V] fp32, past_key_values).

        `past_key_values` is a `DecoderPast`: the inputs of the tokens seen so far.  An incremental call returns the
        logits of the NEW tokens only, computed by one causal pass over cached + new tokens (identical values to a
        K/V-cached step: SURVEY.md finding 4, and the engine's own test_cached_decode_equals_teacher_forced_long);
        the per-token K/V cache that makes decoding O(L) lives inside `generate()` / `cxrm_rollout`."""
        B = self._ensure_cross_kv(pixel_values, encoder_outputs)
        ids = decoder_input_ids.to(self.device)
        q = ids.shape[1]
        past = past_key_values if isinstance(past_key_values, DecoderPast) and past_key_values.length else None
        n_past = past.length if past is not None else 0
        tt = torch.zeros_like(ids) if decoder_token_type_ids is None else decoder_token_type_ids.to(self.device)
        pos = (torch.arange(n_past, n_past + q, device=self.device)[None].expand_as(ids)      # BERT's default positions
               if decoder_position_ids is None else decoder_position_ids.to(self.device))
        # decoder_attention_mask covers past + new tokens (HF semantics, modelling_longitudinal.py:274)
        mask = (torch.ones(ids.shape[0], n_past + q, dtype=torch.int64, device=self.device)
                if decoder_attention_mask is None else decoder_attention_mask.to(self.device))
        if mask.shape[1] != n_past + q:
            raise ValueError(f"decoder_attention_mask has {mask.shape[1]} columns for {n_past} cached + {q} new tokens")
        if past is not None:
            ids_all, tt_all, pos_all = (torch.cat((a, b), dim=1) for a, b in ((past.ids, ids), (past.tt, tt), (past.pos, pos)))
        else:
            ids_all, tt_all, pos_all = ids, tt, pos
        logits = self.engine.decoder_forward(ids_all, tt_all, pos_all, mask, n_studies=B)
        logits = logits[:, n_past:]
        loss = None
        if labels is not None:
            loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1).to(self.device))
        out = ModelOutput(loss=loss, logits=logits)
        if use_cache or past_key_values is not None:
            out["past_key_values"] = DecoderPast(ids_all, tt_all, pos_all)
        return out

    __call__ = forward
