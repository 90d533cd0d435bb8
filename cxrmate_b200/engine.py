"""Thin torch-tensor wrapper over the C ABI: tensors in, tensors out, raw
pointers across the boundary.  PyTorch is only the allocator / stream owner."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import lib as _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class RolloutOutput:
    sequences: torch.Tensor       # [R, P+T] int64 (PAD filled); R = B, or 2B for mode 'both' (sample rows first)
    steps: int                    # executed decode steps (== len(scores) of HF generate)
    logprobs: torch.Tensor        # [R, T] fp32
    margins: torch.Tensor         # [R, T] fp32 decision margins
    topk_idx: torch.Tensor        # [R, T, 64] int32
    topk_val: torch.Tensor        # [R, T, 64] fp32
    topk_cnt: torch.Tensor        # [R, T] int32
    last_logits: torch.Tensor     # [R, V] fp32


@dataclass
class BeamOutput:
    sequences: torch.Tensor       # [B, P + longest hypothesis] int64, PAD filled
    scores: torch.Tensor          # [B] fp32 (HF sequences_scores)
    lengths: torch.Tensor         # [B] int32 generated length of each hypothesis
    steps: int                    # decode steps executed


def normalise_weight_name(name: str) -> str:
    """state_dict key of a real `peft`-wrapped decoder -> the plain key the engine knows (last
    row: `decoder.base_model.model.<...>.base_layer.weight`, `<...>.lora_A.default.weight`).  Keys that peft did not
    touch pass through unchanged; cxrm_finalize_weights rejects anything it cannot place (CXRM_ERR_WEIGHT)."""
    name = name.replace("decoder.base_model.model.", "decoder.")
    name = name.replace(".base_layer.", ".")
    for ab in ("lora_A", "lora_B"):
        name = name.replace(f".{ab}.default.", f".{ab}.")
    return name


class Engine:
    """One engine per process / GPU."""

    def __init__(self, *, dtype: str = "bf16", device: int = 0, image_size: int = 384, max_studies: int = 32,
                 max_images: int = 5, max_prompt: int = 256, max_new_tokens: int = 255, vocab: int = 30000,
                 cvt_depth: Sequence[int] = (1, 4, 16), dec_layers: int = 6, rwd_layers: int = 12,
                 rwd_vocab: int = 30522, rwd_max_len: int = 512, rwd_max_seqs: Optional[int] = None,
                 enc_chunk: int = 64, use_tensor_cores: bool = True, use_cuda_graph: bool = True,
                 max_train_tokens: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("cxrmate_b200.Engine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        cfg = _lib.default_config()
        cfg.dtype = {"fp32": _lib.CXRM_F32, "bf16": _lib.CXRM_BF16}[dtype]
        cfg.image_h = cfg.image_w = image_size
        cfg.max_studies, cfg.max_images = max_studies, max_images
        cfg.max_prompt, cfg.max_new_tokens, cfg.vocab = max_prompt, max_new_tokens, vocab
        for i in range(3):
            cfg.cvt_depth[i] = cvt_depth[i]
        cfg.dec_layers, cfg.rwd_layers, cfg.rwd_vocab, cfg.rwd_max_len = dec_layers, rwd_layers, rwd_vocab, rwd_max_len
        cfg.rwd_max_seqs = rwd_max_seqs if rwd_max_seqs is not None else 3 * max_studies
        cfg.enc_chunk = enc_chunk
        cfg.use_tensor_cores = int(use_tensor_cores)
        cfg.use_cuda_graph = int(use_cuda_graph)
        cfg.max_train_tokens = int(max_train_tokens)
        self.cfg = cfg
        self.dtype = dtype
        self.torch_dtype = torch.float32 if dtype == "fp32" else torch.bfloat16
        self.device = torch.device("cuda", device)
        self.tokens_per_image = (image_size // 16) ** 2
        h = C.c_void_p()
        rc = self.lib.cxrm_create(C.byref(cfg), device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"cxrm_create failed ({rc}): {self.lib.cxrm_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.cxrm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.cxrm_last_error(self.h).decode()}")

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, prefix: str = ""):
        """sd: mapping name -> fp32 tensor (CPU or CUDA) in the reference's naming."""
        for name, t in sd.items():
            if not torch.is_floating_point(t):
                continue                       # num_batches_tracked
            name = normalise_weight_name(name)
            t = t.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            rc = self.lib.cxrm_load_weight(self.h, (prefix + name).encode(), C.c_void_p(t.data_ptr()), shape, t.dim(),
                                           int(t.is_cuda))
            self._check(rc, f"cxrm_load_weight({name})")

    def finalize(self):
        self._check(self.lib.cxrm_finalize_weights(self.h), "cxrm_finalize_weights")

    @property
    def launch_count(self) -> int:
        return int(self.lib.cxrm_launch_count(self.h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.cxrm_workspace_bytes(self.h))

    # ------------------------------------------------------------------ encoder
    def encode(self, pixels: torch.Tensor):
        """pixels [B,N,3,H,W] fp32 cuda -> (memory [B,N*T,768] engine dtype, mask [B,N*T] bool)."""
        assert pixels.is_cuda and pixels.dtype == torch.float32 and pixels.dim() == 5
        pixels = pixels.contiguous()
        B, N = pixels.shape[:2]
        S = N * self.tokens_per_image
        mem = torch.empty(B, S, 768, dtype=self.torch_dtype, device=pixels.device)
        mask = torch.empty(B, S, dtype=torch.uint8, device=pixels.device)
        self._check(self.lib.cxrm_encode(self.h, _ptr(pixels), B, N, _ptr(mem), _ptr(mask), _stream()), "cxrm_encode")
        return mem, mask.bool()

    def prefill_cross_kv(self, memory: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        if memory is None:
            self._check(self.lib.cxrm_prefill_cross_kv(self.h, None, None, 0, 0, _stream()), "cxrm_prefill_cross_kv")
            return
        memory = memory.to(self.torch_dtype).contiguous()
        B, S = memory.shape[:2]
        m8 = None if mask is None else mask.to(torch.uint8).contiguous()
        self._check(self.lib.cxrm_prefill_cross_kv(self.h, _ptr(memory), _ptr(m8), B, S, _stream()),
                    "cxrm_prefill_cross_kv")

    # ------------------------------------------------------------------ rollout
    def rollout(self, prompt_ids: torch.Tensor, *, mode: str, max_new_tokens: int, eos_token_id: int,
                pad_token_id: int, mask_token_id: Optional[int], special_sample=(), sections_sample=(0,),
                special_greedy=(), sections_greedy=(0,), top_k: int = 50, temperature: float = 1.0,
                exp_noise: Optional[torch.Tensor] = None, seed: int = 0, want_margins: bool = True) -> RolloutOutput:
        """want_margins=False skips the decision-margin diagnosis (a second scan of every row per step; the SCST step
        never asks for it) and leaves `margins` unwritten."""
        dev = prompt_ids.device
        B, P = prompt_ids.shape
        p32 = prompt_ids.to(torch.int32).contiguous()
        m = {"greedy": _lib.CXRM_GREEDY, "sample": _lib.CXRM_SAMPLE, "both": _lib.CXRM_BOTH}[mode]
        R = B * (2 if mode == "both" else 1)
        T, V = max_new_tokens, self.cfg.vocab
        out = RolloutOutput(
            sequences=torch.empty(R, P + T, dtype=torch.int32, device=dev), steps=0,
            logprobs=torch.empty(R, T, dtype=torch.float32, device=dev),
            margins=torch.empty(R, T, dtype=torch.float32, device=dev),
            topk_idx=torch.zeros(R, T, _lib.TOPK_CAP, dtype=torch.int32, device=dev),
            topk_val=torch.zeros(R, T, _lib.TOPK_CAP, dtype=torch.float32, device=dev),
            topk_cnt=torch.zeros(R, T, dtype=torch.int32, device=dev),
            last_logits=torch.empty(R, V, dtype=torch.float32, device=dev))
        a = _lib.CxrmRolloutArgs()
        a.mode, a.B, a.P = m, B, P
        a.prompt_ids = p32.data_ptr()
        a.mask_token_id = -1 if mask_token_id is None else int(mask_token_id)
        assert len(sections_sample) == len(special_sample) + 1 and len(sections_greedy) == len(special_greedy) + 1
        a.n_special_sample = len(special_sample)
        a.n_special_greedy = len(special_greedy)
        for i, v in enumerate(special_sample):
            a.special_sample[i] = int(v)
        for i, v in enumerate(sections_sample):
            a.sections_sample[i] = int(v)
        for i, v in enumerate(special_greedy):
            a.special_greedy[i] = int(v)
        for i, v in enumerate(sections_greedy):
            a.sections_greedy[i] = int(v)
        a.max_new_tokens, a.eos_token_id, a.pad_token_id = T, int(eos_token_id), int(pad_token_id)
        a.top_k, a.temperature, a.seed = int(top_k), float(temperature), int(seed)
        if exp_noise is not None:
            assert exp_noise.shape == (T, B, V) and exp_noise.dtype == torch.float32 and exp_noise.is_cuda
            exp_noise = exp_noise.contiguous()
            a.exp_noise = exp_noise.data_ptr()
        a.sequences = out.sequences.data_ptr()
        a.logprobs = out.logprobs.data_ptr()
        a.margins = out.margins.data_ptr() if want_margins else None
        a.topk_idx = out.topk_idx.data_ptr()
        a.topk_val = out.topk_val.data_ptr()
        a.topk_cnt = out.topk_cnt.data_ptr()
        a.last_logits = out.last_logits.data_ptr()
        steps = C.c_int32(0)
        a.steps_out = C.addressof(steps)
        self._check(self.lib.cxrm_rollout(self.h, C.byref(a), _stream()), "cxrm_rollout")
        out.steps = int(steps.value)
        out.sequences = out.sequences.to(torch.int64)
        return out

    def rollout_beam(self, prompt_ids: torch.Tensor, *, num_beams: int, max_new_tokens: int, eos_token_id: int,
                     pad_token_id: int, mask_token_id: Optional[int], special=(), sections=(0,),
                     length_penalty: float = 1.0) -> "BeamOutput":
        """Beam search (cxrm_rollout_beam): prompt [B, P] -> best finished hypothesis per study, trimmed to the longest
        one like HF's `generate(num_beams=...)['sequences']` (without the auto-prepended BOS)."""
        dev = prompt_ids.device
        B, P = prompt_ids.shape
        p32 = prompt_ids.to(torch.int32).contiguous()
        T = max_new_tokens
        seq = torch.empty(B, P + T, dtype=torch.int32, device=dev)
        scores = torch.empty(B, dtype=torch.float32, device=dev)
        lens = torch.empty(B, dtype=torch.int32, device=dev)
        a = _lib.CxrmBeamArgs()
        a.B, a.P, a.prompt_ids = B, P, p32.data_ptr()
        a.mask_token_id = -1 if mask_token_id is None else int(mask_token_id)
        assert len(sections) == len(special) + 1
        a.n_special = len(special)
        for i, v in enumerate(special):
            a.special[i] = int(v)
        for i, v in enumerate(sections):
            a.sections[i] = int(v)
        a.num_beams, a.max_new_tokens = int(num_beams), T
        a.eos_token_id, a.pad_token_id, a.length_penalty = int(eos_token_id), int(pad_token_id), float(length_penalty)
        a.sequences, a.scores, a.lengths = seq.data_ptr(), scores.data_ptr(), lens.data_ptr()
        steps = C.c_int32(0)
        a.steps_out = C.addressof(steps)
        self._check(self.lib.cxrm_rollout_beam(self.h, C.byref(a), _stream()), "cxrm_rollout_beam")
        n = int(lens.max())
        return BeamOutput(sequences=seq[:, : P + n].to(torch.int64), scores=scores, lengths=lens, steps=int(steps.value))

    # ------------------------------------------------------------------ teacher-forced forward
    def decoder_forward(self, ids, token_type_ids, position_ids, key_mask, n_studies: int, last_only: bool = False):
        R, L = ids.shape
        dev = ids.device
        i32 = lambda t: t.to(torch.int32).contiguous()
        ids32, tt32, pos32 = i32(ids), i32(token_type_ids), i32(position_ids)
        km = key_mask.to(torch.uint8).contiguous()
        shape = (R, self.cfg.vocab) if last_only else (R, L, self.cfg.vocab)
        logits = torch.empty(shape, dtype=torch.float32, device=dev)
        self._check(self.lib.cxrm_decoder_forward(self.h, _ptr(ids32), _ptr(tt32), _ptr(pos32), _ptr(km), R, L,
                                                  n_studies, int(last_only), _ptr(logits), _stream()),
                    "cxrm_decoder_forward")
        return logits

    # ------------------------------------------------------------------ reward
    def reward_embed(self, ids: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        n, L = ids.shape
        ids32, lens32 = ids.to(torch.int32).contiguous(), lens.to(torch.int32).contiguous()
        emb = torch.empty(n, 128, dtype=torch.float32, device=ids.device)
        self._check(self.lib.cxrm_reward_embed(self.h, _ptr(ids32), _ptr(lens32), n, L, _ptr(emb), _stream()),
                    "cxrm_reward_embed")
        return emb

    def cosine(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        a, b = a.contiguous().float(), b.contiguous().float()
        out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
        self._check(self.lib.cxrm_cosine(self.h, _ptr(a), _ptr(b), a.shape[0], a.shape[1], _ptr(out), _stream()),
                    "cxrm_cosine")
        return out

    def reinforce_loss(self, logprobs: torch.Tensor, advantage: torch.Tensor) -> torch.Tensor:
        """reference `reinforce_loss` (scst/gen_prompt.py:331-366) from the sample rollout's log-probs [B, T] (0 at PAD)
        and the advantage [B]: mean_b(-sum_t logprob * advantage) -> 0-dim fp32 tensor on the device."""
        assert logprobs.dim() == 2 and advantage.shape == (logprobs.shape[0],)
        lp, adv = logprobs.float(), advantage.contiguous().float()
        assert lp.stride(1) == 1
        out = torch.empty(1, dtype=torch.float32, device=lp.device)
        self._check(self.lib.cxrm_reinforce_loss(self.h, _ptr(lp), lp.stride(0), _ptr(adv), lp.shape[0], lp.shape[1], _ptr(out),
                                                 _stream()), "cxrm_reinforce_loss")
        return out[0]

    def reward(self, pred_ids, pred_lens, label_ids, label_lens) -> torch.Tensor:
        n = pred_ids.shape[0]
        p32, pl = pred_ids.to(torch.int32).contiguous(), pred_lens.to(torch.int32).contiguous()
        l32, ll = label_ids.to(torch.int32).contiguous(), label_lens.to(torch.int32).contiguous()
        out = torch.empty(n, dtype=torch.float32, device=pred_ids.device)
        self._check(self.lib.cxrm_reward(self.h, _ptr(p32), _ptr(pl), p32.shape[1], _ptr(l32), _ptr(ll), l32.shape[1],
                                         n, _ptr(out), _stream()), "cxrm_reward")
        return out

    # ------------------------------------------------------------------ host-buffer SCST step
    def set_id_map(self, id_map: torch.Tensor, cls_id: int, sep_id: int, bos_id: int, sep_dec_id: int,
                   n_special: int = 12):
        """id_map[i] = reward-model id of decoder id i; decoder ids < n_special are special tokens (skipped, like
        tokenizer.decode(skip_special_tokens=True))."""
        m = id_map.to(torch.int32).cpu().contiguous()
        self._check(self.lib.cxrm_set_id_map(self.h, _ptr(m), m.numel(), cls_id, sep_id, bos_id, sep_dec_id, n_special),
                    "cxrm_set_id_map")

    def bridge_ids(self, sequences: torch.Tensor, eos_token_id: int, out_len: int):
        """Device-side text bridge (cxrm_bridge_ids): generated ids [R, L] -> (reward ids [R, out_len], lens [R])."""
        seq = sequences.to(torch.int32).contiguous()
        R, L = seq.shape
        ids = torch.empty(R, out_len, dtype=torch.int32, device=seq.device)
        lens = torch.empty(R, dtype=torch.int32, device=seq.device)
        self._check(self.lib.cxrm_bridge_ids(self.h, _ptr(seq), R, L, int(eos_token_id), _ptr(ids), _ptr(lens), out_len,
                                             _stream()), "cxrm_bridge_ids")
        return ids, lens

    # ------------------------------------------------------------------ training (teacher-forced forward + backward)
    def grad_layout(self, lora_only: bool):
        """[(name, offset, numel, stage)] of the flat fp32 gradient buffer (cxrm_grad_info)."""
        out = []
        name = C.create_string_buffer(256)
        off, ne, st = C.c_int64(), C.c_int64(), C.c_int()
        for i in range(self.lib.cxrm_grad_count(self.h, int(lora_only))):
            self._check(self.lib.cxrm_grad_info(self.h, int(lora_only), i, name, len(name), C.byref(off), C.byref(ne),
                                                C.byref(st)), "cxrm_grad_info")
            out.append((name.value.decode(), int(off.value), int(ne.value), int(st.value)))
        return out

    @property
    def train_stages(self) -> int:
        return int(self.lib.cxrm_train_stages(self.h))

    def train_step(self, ids, token_type_ids, position_ids, key_mask, targets, *, loss_kind: str, ignore_index: int,
                   advantage: Optional[torch.Tensor] = None, top_k: int = 0, temperature: float = 1.0,
                   lora_only: bool = False, grads: Optional[torch.Tensor] = None, stage: int = -1, _keep=None):
        """Teacher-forced decoder forward + backward (cxrm_train_step).  ids / token_type_ids / position_ids / targets
        [R, L] int, key_mask [R, L]; loss_kind 'ce' | 'reinforce'.  Returns (loss 0-dim tensor, flat fp32 gradients).
        stage -1 runs everything; otherwise the caller runs stages 0 .. train_stages-1 in order with the same tensors
        (`_keep` carries the int32 copies between stage calls)."""
        R, L = ids.shape
        if _keep is None:
            i32 = lambda t: t.to(torch.int32).contiguous()
            _keep = dict(ids=i32(ids), tt=i32(token_type_ids), pos=i32(position_ids), km=key_mask.to(torch.uint8).contiguous(),
                         tgt=i32(targets), adv=None if advantage is None else advantage.contiguous().float(),
                         loss=torch.zeros(1, dtype=torch.float32, device=ids.device))
        if grads is None:
            grads = torch.zeros(int(self.lib.cxrm_grad_total(self.h, int(lora_only))), dtype=torch.float32, device=ids.device)
        a = _lib.CxrmTrainArgs()
        a.R, a.L = R, L
        a.ids, a.token_type_ids, a.position_ids = _keep["ids"].data_ptr(), _keep["tt"].data_ptr(), _keep["pos"].data_ptr()
        a.key_mask, a.targets = _keep["km"].data_ptr(), _keep["tgt"].data_ptr()
        a.ignore_index = int(ignore_index)
        a.loss_kind = {"ce": 0, "reinforce": 1}[loss_kind]
        a.advantage = None if _keep["adv"] is None else _keep["adv"].data_ptr()
        a.top_k, a.temperature, a.lora_only = int(top_k), float(temperature), int(lora_only)
        a.loss_out, a.grads = _keep["loss"].data_ptr(), grads.data_ptr()
        self._check(self.lib.cxrm_train_step(self.h, C.byref(a), int(stage), _stream()), "cxrm_train_step")
        self._train_keep = _keep
        return _keep["loss"][0], grads

    def grads_by_name(self, flat: torch.Tensor, lora_only: bool) -> dict:
        return {n: flat[o:o + ne] for n, o, ne, _ in self.grad_layout(lora_only)}

    def last_phase_ms(self) -> dict:
        """device ms of the phases of the last scst_step: encode, cross_kv, rollout, reward, prompt pass (inside rollout)"""
        buf = (C.c_float * 5)()
        self._check(self.lib.cxrm_last_phase_ms(self.h, buf), "cxrm_last_phase_ms")
        return dict(zip(("encode", "cross_kv", "rollout", "reward", "prompt_pass"), (float(x) for x in buf)))

    def set_profile(self, on: bool):
        self._check(self.lib.cxrm_set_profile(self.h, int(on)), "cxrm_set_profile")

    def profile_report(self) -> dict:
        import json
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.cxrm_profile_report(self.h, buf, len(buf)), "cxrm_profile_report")
        return json.loads(buf.value.decode())

    def scst_step(self, pixels: torch.Tensor, prompt_ids: torch.Tensor, label_ids: torch.Tensor,
                  label_lens: torch.Tensor, *, max_new_tokens: int, eos_token_id: int, pad_token_id: int,
                  mask_token_id: int, special_sample, sections_sample, special_greedy, sections_greedy,
                  top_k: int = 50, temperature: float = 1.0, seed: int = 0, out=None,
                  exp_noise: Optional[torch.Tensor] = None):
        """Whole SCST rollout step.  Either every tensor is a HOST tensor (pinned for speed; the timed end-to-end
        call: H2D inside, D2H inside, synchronous) or every tensor is a CUDA tensor (device-resident variant).
        prompt_ids / label_ids / label_lens must be int32."""
        on_dev = pixels.is_cuda
        assert pixels.dtype == torch.float32 and pixels.is_contiguous()
        B, N = pixels.shape[:2]
        P = prompt_ids.shape[1]
        T = max_new_tokens
        p32 = prompt_ids.to(torch.int32).contiguous()
        l32, ll = label_ids.to(torch.int32).contiguous(), label_lens.to(torch.int32).contiguous()
        assert p32.is_cuda == on_dev and l32.is_cuda == on_dev and ll.is_cuda == on_dev
        if out is None:
            pin = dict(device=pixels.device) if on_dev else dict(pin_memory=True)
            out = dict(sequences=torch.empty(2 * B, P + T, dtype=torch.int32, **pin),
                       logprobs=torch.empty(2 * B, T, dtype=torch.float32, **pin),
                       reward=torch.empty(B, dtype=torch.float32, **pin),
                       baseline=torch.empty(B, dtype=torch.float32, **pin),
                       advantage=torch.empty(B, dtype=torch.float32, **pin),
                       steps=torch.zeros(1, dtype=torch.int32, **pin))
        a = _lib.CxrmRolloutArgs()
        a.mask_token_id = int(mask_token_id)
        a.n_special_sample, a.n_special_greedy = len(special_sample), len(special_greedy)
        for i, v in enumerate(special_sample):
            a.special_sample[i] = int(v)
        for i, v in enumerate(sections_sample):
            a.sections_sample[i] = int(v)
        for i, v in enumerate(special_greedy):
            a.special_greedy[i] = int(v)
        for i, v in enumerate(sections_greedy):
            a.sections_greedy[i] = int(v)
        a.max_new_tokens, a.eos_token_id, a.pad_token_id = T, int(eos_token_id), int(pad_token_id)
        a.top_k, a.temperature, a.seed = int(top_k), float(temperature), int(seed)
        if exp_noise is not None:      # validation: the sample rows consume this Exp(1) tensor instead of Philox draws
            assert exp_noise.shape == (T, B, self.cfg.vocab) and exp_noise.dtype == torch.float32 and exp_noise.is_cuda
            exp_noise = exp_noise.contiguous()
            a.exp_noise = exp_noise.data_ptr()
        fn = self.lib.cxrm_scst_step_device if on_dev else self.lib.cxrm_scst_step_host
        rc = fn(
            self.h, _ptr(pixels), B, N, _ptr(p32), P, C.byref(a), _ptr(l32), _ptr(ll), l32.shape[1],
            _ptr(out["sequences"]), _ptr(out["logprobs"]), _ptr(out["reward"]), _ptr(out["baseline"]),
            _ptr(out["advantage"]), _ptr(out["steps"]), _stream())
        self._check(rc, "cxrm_scst_step")
        return out


# ---------------------------------------------------------------------- kernel-level test hooks
def gemm_hook(impl: str, A, W, bias=None, act: int = 0, residual=None, out_f32: bool = False):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    dt = _lib.CXRM_F32 if A.dtype == torch.float32 else _lib.CXRM_BF16
    out = torch.empty(M, N, dtype=torch.float32 if (out_f32 or dt == _lib.CXRM_F32) else torch.bfloat16, device=A.device)
    rc = lib.cxrm_test_gemm({"simt": 0, "tcgen05": 1, "skinny": 2}[impl], dt, _ptr(A), _ptr(W), _ptr(out), M, N, K, _ptr(bias), act,
                            _ptr(residual), int(out_f32), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_gemm failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return out


def sample_hook(logits: torch.Tensor, top_k: int, temperature: float, seed: int, step: int, tmax: int = 256):
    """The sampling head on caller logits [R, V] fp32 (cxrm_test_sample): tokens [R] int32, log-probs [R]."""
    lib = _lib.load()
    logits = logits.contiguous().float()
    R, V = logits.shape
    tok = torch.empty(R, dtype=torch.int32, device=logits.device)
    lp = torch.empty(R, dtype=torch.float32, device=logits.device)
    rc = lib.cxrm_test_sample(_ptr(logits), R, V, int(top_k), float(temperature), int(seed), int(step), int(tmax), _ptr(tok),
                              _ptr(lp), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_sample failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return tok, lp


def gemm_ln_hook(A, W, bias, act, residual, gamma, beta, eps=1e-12):
    """bf16 decode-step GEMM + LayerNorm: skinny split-K GEMM + fused reduce/bias/act/residual/LN pair.  A [M<=64,K], W [N,K]."""
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=A.device)
    ws = torch.empty(8 * 64 * N, dtype=torch.float32, device=A.device)      # up to 8 K-splits of [64, N] fp32
    rc = lib.cxrm_test_gemm_ln(_ptr(A), _ptr(W), _ptr(out), M, N, K, _ptr(bias), act, _ptr(residual), _ptr(gamma),
                               _ptr(beta), float(eps), _ptr(ws), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_gemm_ln failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return out


def attention_hook(q, k, v, key_mask=None, causal: bool = False, scale: float = 0.125):
    """q [b, Lq, h*64], k/v [b, Lk, h*64] -> o like q."""
    lib = _lib.load()
    b, Lq, Cdim = q.shape
    Lk = k.shape[1]
    dt = _lib.CXRM_F32 if q.dtype == torch.float32 else _lib.CXRM_BF16
    o = torch.empty_like(q)
    km = None if key_mask is None else key_mask.to(torch.uint8).contiguous()
    rc = lib.cxrm_test_attention(dt, _ptr(q), _ptr(k), _ptr(v), _ptr(o), b, Cdim // 64, Lq, Lk, _ptr(km), int(causal),
                                 float(scale), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_attention failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return o


def attention_packed_hook(qkv, offsets, lens, heads: int, scale: float = 0.125):
    """qkv [total, 3 * heads*64] (q | k | v per token), sequence i = rows offsets[i] .. + lens[i] -> o [total, heads*64]:
    the packed self-attention of the CXR-BERT reward batch (no mask inside a sequence)."""
    lib = _lib.load()
    total = qkv.shape[0]
    dt = _lib.CXRM_F32 if qkv.dtype == torch.float32 else _lib.CXRM_BF16
    o = torch.zeros(total, heads * 64, dtype=qkv.dtype, device=qkv.device)
    off32, len32 = offsets.to(torch.int32).contiguous(), lens.to(torch.int32).contiguous()
    rc = lib.cxrm_test_attention_packed(dt, _ptr(qkv), _ptr(o), int(lens.numel()), int(heads), int(lens.max()), _ptr(off32),
                                        _ptr(len32), int(total), float(scale), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_attention_packed failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return o


def layernorm_hook(x, gamma, beta, eps: float):
    """x [rows, C] fp32 | bf16 -> LayerNorm(x) like x (the vectorised kernel when rows are 16-byte aligned)."""
    lib = _lib.load()
    dt = _lib.CXRM_F32 if x.dtype == torch.float32 else _lib.CXRM_BF16
    y = torch.empty_like(x)
    rc = lib.cxrm_test_layernorm(dt, _ptr(x), _ptr(y), _ptr(gamma), _ptr(beta), x.shape[0], x.shape[1], float(eps), _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_layernorm failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return y


def ln_dwconv_hook(x, H: int, W: int, cls: int, gamma, beta, eps: float, w, scale, shift):
    """x [n, cls + H*W, C]; w [3, 9, C] fp32 (q, k, v; tap-major); scale / shift [3, C] (folded BatchNorm).
    Returns q [n, cls + H*W, C], k, v [n, cls + Hk*Wk, C]."""
    lib = _lib.load()
    n, _, Cc = x.shape
    Hk, Wk = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    dt = _lib.CXRM_F32 if x.dtype == torch.float32 else _lib.CXRM_BF16
    q = torch.empty_like(x)
    k = torch.empty(n, cls + Hk * Wk, Cc, dtype=x.dtype, device=x.device)
    v = torch.empty_like(k)
    stats = torch.empty(n * (cls + H * W) * 2, dtype=torch.float32, device=x.device)
    rc = lib.cxrm_test_ln_dwconv(dt, _ptr(x), _ptr(q), _ptr(k), _ptr(v), _ptr(stats), _ptr(gamma), _ptr(beta), float(eps),
                                 _ptr(w), _ptr(scale), _ptr(shift), n, H, W, Cc, cls, _stream())
    if rc != 0:
        raise RuntimeError(f"cxrm_test_ln_dwconv failed ({rc}): {lib.cxrm_last_error(None).decode()}")
    return q, k, v
