#!/usr/bin/env python
"""Per-CTA timeline of gemm_tc_wres_kernel (globaltimer stamps, see WTRACE in gemm_tcgen05.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from cxrmate_b200 import lib as L
from cxrmate_b200.engine import gemm_hook
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)
SH = [(18464, 384, 384, 0, False), (18464, 384, 384, 0, True), (18464, 1536, 384, 1, False), (73728, 192, 192, 0, False),
      (73728, 768, 192, 1, False), (294912, 64, 64, 0, True), (294912, 256, 64, 1, False)]
for (M, N, K, act, res) in SH:
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    for _ in range(3):
        gemm_hook("tcgen05", A, W, bias, act, R)
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    lib.cxrm_test_set_gemm_trace(C.c_void_p(buf.data_ptr()))
    gemm_hook("tcgen05", A, W, bias, act, R)
    torch.cuda.synchronize()
    lib.cxrm_test_set_gemm_trace(None)
    t = buf.view(-1, 16).cpu().double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = lambda i: ((t[:, i] - t0) / 1000)
    m = lambda i: f"{rel(i).mean().item():.2f}"
    print(f"M{M} N{N} K{K} act{act} res{int(res)}: {len(t)} CTAs x {t[:,12].mean().item():.1f} tiles | entry {rel(0).max().item():.2f} setup {m(1)} W-resident {m(2)} "
          f"first-A {m(3)} last-MMA {m(4)} first-acc {m(5)} first-stored {m(6)} last-stored {m(7)} exit mean {m(8)} max {rel(8).max().item():.2f} | "
          f"MMA waited: data {t[:,9].mean().item()/1000:.2f} acc {t[:,10].mean().item()/1000:.2f}; epilogue waited for acc {t[:,11].mean().item()/1000:.2f} us")

if os.environ.get("KBTRACE"):
    # debug build (CXRM_DEFINES=CXRM_WRES_KBTRACE): clocks of the k-blocks of tile 1
    for (M, N, K, act, res) in [(18464, 384, 384, 0, False), (18464, 1536, 384, 1, False)]:
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        for _ in range(3):
            gemm_hook("tcgen05", A, W, bias, act, None)
        buf = torch.zeros(148 * 16 + 148 * 32, dtype=torch.int64, device="cuda")
        lib.cxrm_test_set_gemm_trace(C.c_void_p(buf.data_ptr()))
        gemm_hook("tcgen05", A, W, bias, act, None)
        torch.cuda.synchronize()
        lib.cxrm_test_set_gemm_trace(None)
        q = buf[148 * 16:].view(148, 8, 4).cpu()
        for cta in (0, 1, 77):
            base = q[cta, 0, 0].item()
            print(f"M{M} N{N} cta {cta}:", " | ".join(f"kb{kb} start {q[cta,kb,0].item()-base} mma-issued +{q[cta,kb,1].item()-q[cta,kb,0].item()} commit +{q[cta,kb,2].item()-q[cta,kb,1].item()}" for kb in range(6)))
