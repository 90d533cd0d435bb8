#!/usr/bin/env python
"""Steady-state cost of the decode-step GEMM kernels: N back-to-back launches captured in a CUDA graph.
    python tools/microbench_decode.py   (on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cxrmate_b200 import lib as L
from cxrmate_b200.engine import gemm_hook, gemm_ln_hook

lib = L.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def bench(name, fn, n=60, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            for _ in range(n):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(reps):
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1000 / n)
    print(f"{name:50s} {best:8.2f} us per call", flush=True)


def mk(M, N, K):
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    W = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    res = torch.randn(M, N, device=dev, generator=g).bfloat16()
    return A, W, bias, res


for pdl in (0, 1):
    lib.cxrm_test_set_pdl(pdl)
    print(f"--- pdl={pdl}")
    for (M, N, K) in [(64, 768, 768), (64, 768, 3072)]:
        A, W, bias, res = mk(M, N, K)
        bench(f"skinny direct {M}x{N}x{K}", lambda: gemm_hook("skinny", A, W, bias, 0, None, N > 8192))
        if N <= 1024:
            gm, bt = torch.ones(N, device=dev), torch.zeros(N, device=dev)
            bench(f"skinny split-K + LN {M}x{N}x{K}", lambda: gemm_ln_hook(A, W, bias, 0, res, gm, bt), n=30)
        bench(f"generic tcgen05 {M}x{N}x{K}", lambda: gemm_hook("tcgen05", A, W, bias, 0, None, N > 8192))
lib.cxrm_test_set_pdl(0)
x = torch.zeros(64, 768, device=dev)
bench("torch elementwise add 64x768 (launch floor)", lambda: x.add_(1.0))
