"""Kernel-level parity on the GPU, through the C ABI test hooks: each CUDA kernel
against a plain PyTorch fp32 evaluation of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_gemm(A, W, bias, act, residual):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


SHAPES = [
    (128, 64, 64), (300, 192, 576), (1000, 384, 1728), (577, 1536, 384), (64, 2304, 768), (64, 768, 3072),
    (33, 30000, 768), (4096, 256, 64), (129, 100, 152), (2, 128, 768), (700, 768, 768),
    # >= 2 tiles per SM: the persistent kernel (double-buffered TMEM accumulators), tile widths 128 / 64 / 256 / 96
    (20000, 384, 384), (38000, 64, 64), (19000, 768, 1536), (37900, 96, 192),
    # short K, huge M: the weight-resident persistent kernel (column slices of 128 / 192 / 256 / 64 / 128, ragged last row tile)
    (18464, 1536, 384), (40000, 192, 192), (40010, 768, 192), (60000, 64, 152), (40000, 256, 64), (18432, 768, 384),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("impl,dtype", [("simt", torch.float32), ("simt", torch.bfloat16), ("tcgen05", torch.bfloat16)])
def test_gemm(impl, dtype, M, N, K):
    from cxrmate_b200.engine import gemm_hook
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    for use_bias, act, use_res, out_f32 in [(False, 0, False, False), (True, 1, False, False), (True, 0, True, False),
                                            (True, 0, False, True)]:
        out = gemm_hook(impl, A, W, bias if use_bias else None, act, res if use_res else None, out_f32)
        torch.cuda.synchronize()
        ref = _ref_gemm(A, W, bias if use_bias else None, act, res if use_res else None)
        err = (out.float() - ref).abs().max().item()
        # bf16 stores: half an ulp of the largest output (2^-8 relative) on top of the accumulation error
        tol = 2e-4 if dtype == torch.float32 else (2e-3 if out_f32 else 2e-3 + ref.abs().max().item() * 2 ** -8)
        assert err < tol, f"{impl} {dtype} {M}x{N}x{K} bias={use_bias} act={act} res={use_res} f32={out_f32}: max err {err}"


def test_gemm_tcgen05_matches_simt_bf16():
    """same bf16 inputs, fp32 outputs: tensor-core and SIMT paths differ only by accumulation order"""
    from cxrmate_b200.engine import gemm_hook
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(513, 768, device="cuda", generator=g).bfloat16()
    W = (torch.randn(777, 768, device="cuda", generator=g) * 0.04).bfloat16()
    a = gemm_hook("tcgen05", A, W, out_f32=True)
    b = gemm_hook("simt", A, W, out_f32=True)
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() < 1e-3


SKINNY = [(64, 2304, 768), (64, 768, 768), (64, 768, 3072), (64, 3072, 768), (64, 30000, 768), (33, 30000, 768),
          (2, 128, 768), (7, 768, 3072), (64, 100, 152), (1, 768, 64)]


@pytest.mark.parametrize("M,N,K", SKINNY)
def test_gemm_skinny(M, N, K):
    """decode-step GEMM (M <= 64): deep TMA ring, 64-row A boxes, BN 16/32/64"""
    from cxrmate_b200.engine import gemm_hook
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    for use_bias, act, use_res, out_f32 in [(False, 0, False, False), (True, 1, False, False), (True, 0, True, False),
                                            (True, 0, False, True)]:
        out = gemm_hook("skinny", A, W, bias if use_bias else None, act, res if use_res else None, out_f32)
        torch.cuda.synchronize()
        ref = _ref_gemm(A, W, bias if use_bias else None, act, res if use_res else None)
        err = (out.float() - ref).abs().max().item()
        tol = 2e-3 if out_f32 else 2e-3 + ref.abs().max().item() * 2 ** -8
        assert err < tol, f"skinny {M}x{N}x{K} bias={use_bias} act={act} res={use_res} f32={out_f32}: max err {err}"


@pytest.mark.parametrize("M,N,K", [(64, 768, 768), (64, 768, 3072), (5, 768, 768), (64, 128, 768), (33, 1024, 512)])
@pytest.mark.parametrize("act,use_res", [(0, True), (1, False)])
def test_gemm_splitk_layernorm(M, N, K, act, use_res):
    """skinny split-K GEMM -> fused reduce + bias + act + residual + LayerNorm(eps 1e-12)"""
    from cxrmate_b200.engine import gemm_ln_hook
    g = torch.Generator(device="cuda").manual_seed(M + N + K + act)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if use_res else None
    gamma = 1 + 0.1 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    out = gemm_ln_hook(A, W, bias, act, res, gamma, beta)
    torch.cuda.synchronize()
    pre = _ref_gemm(A, W, bias, act, res).bfloat16().float()      # the GEMM result is stored as bf16 before the LN
    ref = torch.nn.functional.layer_norm(pre, (N,), gamma, beta, 1e-12)
    err = (out.float() - ref).abs().max().item()
    assert err < 4e-2, f"gemm_ln {M}x{N}x{K} act={act} res={use_res}: max err {err}"   # one bf16 ulp at |y| ~ 4


def _ref_attn(q, k, v, heads, key_mask, causal, scale):
    b, Lq, C = q.shape
    Lk = k.shape[1]
    qh = q.float().view(b, Lq, heads, 64).transpose(1, 2)
    kh = k.float().view(b, Lk, heads, 64).transpose(1, 2)
    vh = v.float().view(b, Lk, heads, 64).transpose(1, 2)
    s = qh @ kh.transpose(2, 3) * scale
    neg = torch.finfo(torch.float32).min
    if key_mask is not None:
        s = s.masked_fill(~key_mask.bool()[:, None, None, :], neg)
    if causal:
        qi = torch.arange(Lq, device=q.device)[:, None] + (Lk - Lq)
        kj = torch.arange(Lk, device=q.device)[None, :]
        s = s.masked_fill((kj > qi)[None, None], neg)
    o = torch.softmax(s, -1) @ vh
    return o.transpose(1, 2).reshape(b, Lq, C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,h,Lq,Lk,masked,causal", [
    (2, 1, 256, 64, False, False), (3, 6, 577, 145, False, False), (2, 12, 37, 37, True, True),
    (2, 12, 200, 200, True, False), (1, 3, 130, 70, False, False), (4, 12, 5, 5, True, True),
    # dense multi-tile cases of the tcgen05 kernel: CvT stage 1 / stage 2 shapes, ragged key and query tails
    (1, 1, 2304, 576, False, False), (2, 3, 576, 144, False, False), (2, 2, 300, 401, False, False),
    (1, 1, 64, 16, False, False), (3, 6, 129, 257, False, False),
])
def test_attention(dtype, b, h, Lq, Lk, masked, causal):
    from cxrmate_b200.engine import attention_hook
    g = torch.Generator(device="cuda").manual_seed(Lq * 13 + Lk)
    q = torch.randn(b, Lq, h * 64, device="cuda", generator=g).to(dtype)
    k = torch.randn(b, Lk, h * 64, device="cuda", generator=g).to(dtype)
    v = torch.randn(b, Lk, h * 64, device="cuda", generator=g).to(dtype)
    km = None
    if masked:
        km = torch.rand(b, Lk, device="cuda", generator=g) > 0.3
        km[:, 0] = True
    scale = 0.125
    o = attention_hook(q, k, v, km, causal, scale)
    torch.cuda.synchronize()
    ref = _ref_attn(q, k, v, h, km, causal, scale)
    err = (o.float() - ref).abs().max().item()
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert err < tol, f"attention {dtype} b={b} h={h} {Lq}x{Lk} masked={masked} causal={causal}: max err {err}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("lens,heads", [([257, 64, 130, 1, 200, 192, 193], 12), ([100] * 5, 3), ([300, 17], 2)])
def test_attention_packed(dtype, lens, heads):
    """packed (ragged) self-attention of the reward batch: bf16 runs on the tcgen05 kernel (sequences of 64+ rows in the
    longest), rows of one sequence must not see its neighbours' keys"""
    from cxrmate_b200.engine import attention_packed_hook
    g = torch.Generator(device="cuda").manual_seed(sum(lens) + heads)
    C = heads * 64
    total = sum(lens)
    qkv = torch.randn(total, 3 * C, device="cuda", generator=g).to(dtype)
    lens_t = torch.tensor(lens, device="cuda")
    offs = torch.cumsum(lens_t, 0) - lens_t
    o = attention_packed_hook(qkv, offs, lens_t, heads, 0.125)
    torch.cuda.synchronize()
    ref = torch.empty(total, C, device="cuda")
    for off, n in zip(offs.tolist(), lens):
        blk = qkv[off:off + n].float()
        ref[off:off + n] = _ref_attn(blk[None, :, :C], blk[None, :, C:2 * C], blk[None, :, 2 * C:], heads, None, False, 0.125)[0]
    err = (o.float() - ref).abs().max().item()
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert err < tol, f"packed attention {dtype} lens={lens} heads={heads}: max err {err}"


# ------------------------------------------------------------------------------ LayerNorm / CvT attention front end
@pytest.mark.parametrize("rows,C", [(1000, 64), (577, 192), (333, 384), (64, 768), (50, 128), (7, 100)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm(rows, C, dtype):
    """vectorised LayerNorm (2..32 lanes per row) and the scalar fallback (C = 100) vs torch, eps of CvT and of BERT"""
    from cxrmate_b200.engine import layernorm_hook
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 3 + 1.5).to(dtype)
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    for eps in (1e-5, 1e-12):
        y = layernorm_hook(x, gamma, beta, eps)
        torch.cuda.synchronize()
        ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, eps)
        err = (y.float() - ref).abs().max().item()
        tol = 2e-5 if dtype == torch.float32 else ref.abs().max().item() * 2 ** -8 + 1e-3
        assert err < tol, f"layernorm {rows}x{C} {dtype} eps {eps}: max err {err}"


@pytest.mark.parametrize("n,H,W,C,cls", [(2, 16, 16, 64, 0), (3, 8, 8, 192, 0), (2, 4, 4, 384, 1), (1, 24, 24, 384, 1),
                                         (2, 3, 5, 64, 0), (1, 7, 7, 192, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ln_dwconv_qkv(n, H, W, C, cls, dtype):
    """fused LayerNorm -> depth-wise 3x3 (+ folded BatchNorm) q/k/v projections vs torch (HF modeling_cvt.py:124-141,
    215-228): even, odd and non-square maps, with and without the cls token (which bypasses the convolution)."""
    from cxrmate_b200.engine import ln_dwconv_hook
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + H * 10 + C)
    x = torch.randn(n, cls + H * W, C, device="cuda", generator=g).to(dtype)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    wconv = 0.3 * torch.randn(3, C, 1, 3, 3, device="cuda", generator=g)            # q, k, v depth-wise kernels
    scale = 1 + 0.1 * torch.randn(3, C, device="cuda", generator=g)
    shift = 0.1 * torch.randn(3, C, device="cuda", generator=g)
    w = wconv.reshape(3, C, 9).permute(0, 2, 1).contiguous()                       # [3][9][C]
    q, k, v = ln_dwconv_hook(x, H, W, cls, gamma, beta, 1e-5, w, scale, shift)
    torch.cuda.synchronize()
    y = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    ycls, ymap = y[:, :cls], y[:, cls:].transpose(1, 2).reshape(n, C, H, W)
    outs = []
    for i, stride in enumerate((1, 2, 2)):
        o = torch.nn.functional.conv2d(ymap, wconv[i], stride=stride, padding=1, groups=C)
        o = o * scale[i][None, :, None, None] + shift[i][None, :, None, None]
        outs.append(torch.cat((ycls, o.flatten(2).transpose(1, 2)), 1))
    for name, got, ref in zip("qkv", (q, k, v), outs):
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        err = (got.float() - ref).abs().max().item()
        tol = 1e-4 if dtype == torch.float32 else ref.abs().max().item() * 2 ** -7 + 5e-3
        assert err < tol, f"{name} {n}x{H}x{W}x{C} cls={cls} {dtype}: max err {err}"
