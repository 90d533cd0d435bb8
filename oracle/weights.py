"""Synthetic weights for the oracle: the generator itself is an input generator shared with the product side
(cxrmate_b200/synthetic_weights.py) so that both sides of every parity test see the same tensors."""
from cxrmate_b200.synthetic_weights import *  # noqa: F401,F403
from cxrmate_b200.synthetic_weights import (  # noqa: F401
    CVT_DEPTH, CVT_EMBED_DIM, CVT_HEADS, CVT_PAD, CVT_PATCH, CVT_STRIDE, LORA_ALPHA, LORA_R, count_params,
    make_cxrbert_weights, make_cxrmate_weights,
)
