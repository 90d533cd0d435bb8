"""ctypes binding of libcxrm.so (include/cxrm.h).  No torch types cross the boundary:
only raw device/host pointers, sizes and the stream handle.

The library is mandatory: there is no Python/PyTorch fallback for any entry
point.  Importing this module without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcxrm.so")

CXRM_F32, CXRM_BF16 = 0, 1
CXRM_GREEDY, CXRM_SAMPLE, CXRM_BOTH = 1, 2, 3
TOPK_CAP = 64


class CxrmConfig(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("image_h", C.c_int), ("image_w", C.c_int), ("max_studies", C.c_int),
        ("max_images", C.c_int), ("max_prompt", C.c_int), ("max_new_tokens", C.c_int), ("vocab", C.c_int),
        ("cvt_depth", C.c_int * 3), ("dec_layers", C.c_int), ("rwd_layers", C.c_int), ("rwd_vocab", C.c_int),
        ("rwd_max_len", C.c_int), ("rwd_max_seqs", C.c_int), ("enc_chunk", C.c_int),
        ("use_tensor_cores", C.c_int), ("use_cuda_graph", C.c_int), ("max_train_tokens", C.c_int),
    ]


class CxrmRolloutArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("B", C.c_int), ("P", C.c_int), ("prompt_ids", C.c_void_p), ("mask_token_id", C.c_int),
        ("n_special_sample", C.c_int), ("special_sample", C.c_int * 8), ("sections_sample", C.c_int * 9),
        ("n_special_greedy", C.c_int), ("special_greedy", C.c_int * 8), ("sections_greedy", C.c_int * 9),
        ("max_new_tokens", C.c_int), ("eos_token_id", C.c_int), ("pad_token_id", C.c_int), ("top_k", C.c_int),
        ("temperature", C.c_float), ("exp_noise", C.c_void_p), ("seed", C.c_uint64),
        ("sequences", C.c_void_p), ("logprobs", C.c_void_p), ("margins", C.c_void_p), ("topk_idx", C.c_void_p),
        ("topk_val", C.c_void_p), ("topk_cnt", C.c_void_p), ("last_logits", C.c_void_p), ("steps_out", C.c_void_p),
    ]


class CxrmBeamArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("P", C.c_int), ("prompt_ids", C.c_void_p), ("mask_token_id", C.c_int),
        ("n_special", C.c_int), ("special", C.c_int * 8), ("sections", C.c_int * 9),
        ("num_beams", C.c_int), ("max_new_tokens", C.c_int), ("eos_token_id", C.c_int), ("pad_token_id", C.c_int),
        ("length_penalty", C.c_float), ("sequences", C.c_void_p), ("scores", C.c_void_p), ("lengths", C.c_void_p),
        ("steps_out", C.c_void_p),
    ]


class CxrmTrainArgs(C.Structure):
    _fields_ = [
        ("R", C.c_int), ("L", C.c_int), ("ids", C.c_void_p), ("token_type_ids", C.c_void_p), ("position_ids", C.c_void_p),
        ("key_mask", C.c_void_p), ("targets", C.c_void_p), ("ignore_index", C.c_int), ("loss_kind", C.c_int),
        ("advantage", C.c_void_p), ("top_k", C.c_int), ("temperature", C.c_float), ("lora_only", C.c_int),
        ("loss_out", C.c_void_p), ("grads", C.c_void_p),
    ]


# every symbol include/cxrm.h declares; tests/test_abi.py checks the list against the header
SYMBOLS = {
    "cxrm_default_config": (None, [C.POINTER(CxrmConfig)]),
    "cxrm_create": (C.c_int, [C.POINTER(CxrmConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "cxrm_destroy": (None, [C.c_void_p]),
    "cxrm_last_error": (C.c_char_p, [C.c_void_p]),
    "cxrm_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "cxrm_launch_count": (C.c_uint64, [C.c_void_p]),
    "cxrm_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "cxrm_finalize_weights": (C.c_int, [C.c_void_p]),
    "cxrm_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cxrm_prefill_cross_kv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cxrm_rollout": (C.c_int, [C.c_void_p, C.POINTER(CxrmRolloutArgs), C.c_void_p]),
    "cxrm_preprocess_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "cxrm_rollout_beam": (C.c_int, [C.c_void_p, C.POINTER(CxrmBeamArgs), C.c_void_p]),
    "cxrm_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cxrm_reward_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cxrm_cosine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cxrm_reinforce_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "cxrm_reward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p]),
    "cxrm_set_id_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cxrm_bridge_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p]),
    "cxrm_test_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "cxrm_scst_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                      C.POINTER(CxrmRolloutArgs), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cxrm_scst_step_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.POINTER(CxrmRolloutArgs), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cxrm_last_phase_ms": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cxrm_train_step": (C.c_int, [C.c_void_p, C.POINTER(CxrmTrainArgs), C.c_int, C.c_void_p]),
    "cxrm_train_stages": (C.c_int, [C.c_void_p]),
    "cxrm_grad_count": (C.c_int, [C.c_void_p, C.c_int]),
    "cxrm_grad_total": (C.c_int64, [C.c_void_p, C.c_int]),
    "cxrm_grad_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "cxrm_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "cxrm_profile_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "cxrm_test_gemm": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "cxrm_test_gemm_ln": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "cxrm_test_set_pdl": (None, [C.c_int]),
    "cxrm_test_set_gemm_trace": (None, [C.c_void_p]),
    "cxrm_test_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "cxrm_test_attention_packed": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_longlong, C.c_float, C.c_void_p]),
    "cxrm_test_layernorm": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                      C.c_float, C.c_void_p]),
    "cxrm_test_ln_dwconv": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libcxrm.so and type every entry point.  Raises when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m cxrmate_b200.build` "
            "(the engine is CUDA-only; there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError when the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def default_config() -> CxrmConfig:
    cfg = CxrmConfig()
    load().cxrm_default_config(C.byref(cfg))
    return cfg
