"""CvT-21 + projection head, restated functionally from a state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  CPU, fp32, eval-mode semantics.

Follows:
* HF transformers 5.5.0 `models/cvt/modeling_cvt.py`
    - CvtConvEmbeddings.forward        :99-121   (conv -> tokens -> LayerNorm(eps 1e-5))
    - CvtSelfAttentionConvProjection   :124-141  (depth-wise 3x3, no bias, BatchNorm2d eval)
    - CvtSelfAttention.forward         :211-244  (cls token bypasses the conv; scale = embed_dim**-0.5)
    - CvtLayer.forward                 :371-390  (pre-LN, two residuals)
    - CvtStage.forward                 :436-453  (cls token only in stage 3, dropped on exit)
* reference `modules/transformers/longitudinal_model/modelling_longitudinal.py`
    - CvtProjectionHead.forward        :40-43    (LayerNorm eps 1e-12 -> Linear no bias)
    - MultiCvtWithProjectionHead.forward :56-90  (flatten studies, regroup, mask from pixel[...,0,0,0] != 0)
* reference `modules/transformers/single_model/modelling_single.py:53-78` (no mask).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .weights import CVT_DEPTH, CVT_EMBED_DIM, CVT_HEADS, CVT_PAD, CVT_PATCH, CVT_STRIDE

LN_EPS_CVT = 1e-5      # nn.LayerNorm default inside CvT (modeling_cvt.py:112,366-367)
BN_EPS = 1e-5          # nn.BatchNorm2d default (modeling_cvt.py:135)
LN_EPS_HEAD = 1e-12    # config.layer_norm_eps (modelling_longitudinal.py:34)


def _dw_bn(sd, prefix, img, stride):
    y = F.conv2d(img, sd[prefix + "convolution.weight"], None, stride=stride, padding=1, groups=img.shape[1])
    return F.batch_norm(
        y,
        sd[prefix + "normalization.running_mean"],
        sd[prefix + "normalization.running_var"],
        sd[prefix + "normalization.weight"],
        sd[prefix + "normalization.bias"],
        training=False,
        eps=BN_EPS,
    )


def cvt_layer(sd, p, tokens, H, W, C, heads, with_cls):
    n = tokens.shape[0]
    y = F.layer_norm(tokens, (C,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], LN_EPS_CVT)
    if with_cls:
        cls, y = y[:, :1], y[:, 1:]
    img = y.transpose(1, 2).reshape(n, C, H, W)
    a = p + "attention.attention."
    qkv = []
    for nm, stride in (("query", 1), ("key", 2), ("value", 2)):
        t = _dw_bn(sd, a + f"convolution_projection_{nm}.convolution_projection.", img, stride)
        t = t.flatten(2).transpose(1, 2)
        if with_cls:
            t = torch.cat((cls, t), dim=1)
        t = F.linear(t, sd[a + f"projection_{nm}.weight"], sd[a + f"projection_{nm}.bias"])
        qkv.append(t.reshape(n, t.shape[1], heads, C // heads).permute(0, 2, 1, 3))
    q, k, v = qkv
    score = torch.einsum("bhlk,bhtk->bhlt", q, k) * (C ** -0.5)      # embed_dim, not head_dim (:183)
    prob = torch.softmax(score, dim=-1)
    ctx = torch.einsum("bhlt,bhtv->bhlv", prob, v).permute(0, 2, 1, 3).reshape(n, -1, C)
    attn = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    tokens = attn + tokens
    z = F.layer_norm(tokens, (C,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], LN_EPS_CVT)
    z = F.gelu(F.linear(z, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    z = F.linear(z, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return z + tokens


def cvt_stage_embed(sd, s, x):
    p = f"encoder.cvt.encoder.stages.{s}.embedding.convolution_embeddings."
    C = CVT_EMBED_DIM[s]
    x = F.conv2d(x, sd[p + "projection.weight"], sd[p + "projection.bias"], stride=CVT_STRIDE[s], padding=CVT_PAD[s])
    n, _, H, W = x.shape
    tokens = x.flatten(2).transpose(1, 2)
    tokens = F.layer_norm(tokens, (C,), sd[p + "normalization.weight"], sd[p + "normalization.bias"], LN_EPS_CVT)
    return tokens, H, W


def cvt_features(sd, pixel_values, depth=CVT_DEPTH, return_stages=False):
    """pixel_values [n,3,Himg,Wimg] -> tokens [n, H3*W3, 384] (CvtModel.last_hidden_state, token-major)."""
    x = pixel_values
    stages = []
    for s, C in enumerate(CVT_EMBED_DIM):
        tokens, H, W = cvt_stage_embed(sd, s, x)
        n = tokens.shape[0]
        with_cls = s == 2
        if with_cls:
            cls = sd[f"encoder.cvt.encoder.stages.{s}.cls_token"].expand(n, -1, -1)
            tokens = torch.cat((cls, tokens), dim=1)
        for i in range(depth[s]):
            tokens = cvt_layer(sd, f"encoder.cvt.encoder.stages.{s}.layers.{i}.", tokens, H, W, C, CVT_HEADS[s], with_cls)
        if with_cls:
            tokens = tokens[:, 1:]
        stages.append(tokens)
        x = tokens.transpose(1, 2).reshape(n, C, H, W)
    if return_stages:
        return stages
    return tokens


def projection_head(sd, tokens):
    C = tokens.shape[-1]
    y = F.layer_norm(tokens, (C,), sd["encoder.projection_head.layer_norm.weight"],
                     sd["encoder.projection_head.layer_norm.bias"], LN_EPS_HEAD)
    return F.linear(y, sd["encoder.projection_head.projection.weight"])


def encode_multi(sd, pixel_values, depth=CVT_DEPTH):
    """MultiCvtWithProjectionHead.forward: pixel_values [B,N,3,H,W] ->
    (last_hidden_state [B, N*T, 768], attention_mask [B, N*T] bool)."""
    B, N = pixel_values.shape[:2]
    tokens = cvt_features(sd, pixel_values.reshape(-1, *pixel_values.shape[2:]), depth)
    proj = projection_head(sd, tokens)
    T = proj.shape[1]
    proj = proj.reshape(B, N * T, proj.shape[-1])
    mask = (pixel_values[:, :, 0, 0, 0] != 0.0).repeat_interleave(T, dim=1)
    return proj, mask


def encode_single(sd, pixel_values, depth=CVT_DEPTH):
    """CvtWithProjectionHead.forward: pixel_values [B,3,H,W] -> [B, T, 768]; no mask."""
    return projection_head(sd, cvt_features(sd, pixel_values, depth))
