import sys, os
sys.path.insert(0, "/root/repo")
import torch
from cxrmate_b200.engine import gemm_ln_hook
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 64, 768, 768
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda", generator=g); res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
gm, bt = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
for _ in range(6):
    gemm_ln_hook(A, W, bias, 0, res, gm, bt, cluster=True)
    gemm_ln_hook(A, W, bias, 0, res, gm, bt, cluster=False)
torch.cuda.synchronize()
