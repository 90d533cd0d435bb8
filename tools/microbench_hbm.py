#!/usr/bin/env python
"""Pure-read and copy bandwidth on this GPU (context for the decode-attention roofline)."""
import torch
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for mb in (157, 1024, 4096):
    x = torch.empty(mb * 1024 * 1024 // 2, dtype=torch.bfloat16, device="cuda").normal_()
    y = torch.empty_like(x)
    xf = x.view(torch.float32)
    ms = t(lambda: y.copy_(x)); print(f"{mb} MB copy      : {2 * mb * 1.048576 / ms:.1f} GB/s (read+write)")
    ms = t(lambda: torch.sum(xf)); print(f"{mb} MB sum(fp32) : {mb * 1.048576 / ms:.1f} GB/s (read)")
    ms = t(lambda: torch.amax(x)); print(f"{mb} MB amax(bf16): {mb * 1.048576 / ms:.1f} GB/s (read)")
