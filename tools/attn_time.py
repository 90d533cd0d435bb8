"""Time the dense encoder attention calls (CvT stage shapes, 100 images) through the test hook.
CXRM_NO_TC5_ATTN=1 selects the mma.sync kernel for an A/B.  Usage: python tools/attn_time.py [n_images]"""
import sys
import torch
from cxrmate_b200.engine import attention_hook

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
shapes = [("stage1", n, 1, 9216, 2304, 64 ** -0.5, 1), ("stage2", n, 3, 2304, 576, 192 ** -0.5, 4),
          ("stage3", n, 6, 577, 145, 384 ** -0.5, 16)]
total = 0.0
for name, b, h, Lq, Lk, scale, layers in shapes:
    q = torch.randn(b, Lq, h * 64, device="cuda").bfloat16()
    k = torch.randn(b, Lk, h * 64, device="cuda").bfloat16()
    v = torch.randn(b, Lk, h * 64, device="cuda").bfloat16()
    for _ in range(3):
        attention_hook(q, k, v, None, False, scale)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        attention_hook(q, k, v, None, False, scale)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 4.0 * b * h * Lq * Lk * 64 / ms / 1e9
    total += ms * layers
    print(f"{name}: {ms:.3f} ms/call  {tf:.0f} TF/s  x{layers} layers = {ms * layers:.3f} ms")
print(f"encoder attention total {total:.3f} ms")
