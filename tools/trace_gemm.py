#!/usr/bin/env python
"""Phase timeline of gemm_tc_kernel CTAs (globaltimer stamps): where a tile's time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from cxrmate_b200 import lib as L
from cxrmate_b200.engine import gemm_hook
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K, act, res) in [(18464, 1536, 384, 0, False), (18464, 384, 384, 0, True), (18464, 1536, 384, 1, False), (18464, 384, 1536, 0, True), (24672, 3072, 768, 1, False), (24672, 768, 3072, 0, True)]:
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    for _ in range(3):
        gemm_hook("tcgen05", A, W, bias, act, R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gemm_hook("tcgen05", A, W, bias, act, R); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000
    nct = 8 * 4096 * 8
    buf = torch.zeros(nct, dtype=torch.int64, device="cuda")
    lib.cxrm_test_set_gemm_trace(C.c_void_p(buf.data_ptr()))
    gemm_hook("tcgen05", A, W, bias, act, R)
    torch.cuda.synchronize()
    lib.cxrm_test_set_gemm_trace(None)
    t = buf.view(-1, 8).cpu().double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    d = lambda a, b: (t[:, b] - t[:, a]).mean().item() / 1000
    print(f"M{M} N{N} K{K} act{act} res{int(res)}: {us:.1f} us, {2*M*N*K/us/1e6:.0f} TF/s, {len(t)} CTAs, span {(t[:,7].max()-t0)/1000:.1f} us | "
          f"setup {d(0,1):.2f} first-data {d(1,2):.2f} mainloop {d(2,3):.2f} epi-wait {d(4,5):.2f} epilogue {d(5,6):.2f} total {d(0,7):.2f} us")
