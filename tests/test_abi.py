"""The C-ABI library builds, loads, and exports every symbol include/cxrm.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cxrm.h")).read()
    return sorted(set(re.findall(r"CXRM_API\s+[\w\s\*]+?\b(cxrm_\w+)\s*\(", src)))


def test_header_declares_the_path_entry_points():
    syms = header_symbols()
    for s in ("cxrm_create", "cxrm_load_weight", "cxrm_encode", "cxrm_prefill_cross_kv", "cxrm_rollout",
              "cxrm_decoder_forward", "cxrm_reward", "cxrm_scst_step_host"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from cxrmate_b200 import lib
    assert os.path.exists(lib.LIB_PATH), "run `python -m cxrmate_b200.build` (or __graft_entry__.build())"
    dll = ctypes.CDLL(lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(dll, s), f"{s} is declared in include/cxrm.h but not exported by libcxrm.so"
    assert sorted(lib.SYMBOLS) == header_symbols(), "cxrmate_b200/lib.py binds a different symbol set than the header"


def test_struct_layouts_match_the_header():
    from cxrmate_b200 import lib
    cfg = lib.default_config()
    assert (cfg.image_h, cfg.max_studies, cfg.max_images, cfg.max_prompt, cfg.max_new_tokens) == (384, 32, 5, 256, 255)
    assert list(cfg.cvt_depth) == [1, 4, 16] and cfg.vocab == 30000 and cfg.dec_layers == 6 and cfg.rwd_layers == 12
    assert cfg.use_cuda_graph == 1 and cfg.use_tensor_cores == 1   # last fields: the whole struct lines up
    # pointer members are 8-byte aligned exactly like the C struct
    assert lib.CxrmRolloutArgs.prompt_ids.offset == 16
    assert ctypes.sizeof(lib.CxrmConfig) == 20 * 4   # 17 scalar ints + cvt_depth[3]


def test_no_cpu_fallback():
    """without a CUDA device the product path must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cxrmate_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()
    from cxrmate_b200 import lib
    h = ctypes.c_void_p()
    cfg = lib.default_config()
    rc = lib.load().cxrm_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.load().cxrm_last_error(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cxrmate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports the oracle"
