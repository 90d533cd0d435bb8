"""Time the encoder's GEMM shapes through the test hook (CUDA events, L2 flushed between launches).
    python tools/bench_gemm.py            # run twice: default and CXRM_NO_WRES_GEMM=1"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cxrmate_b200.engine import gemm_hook  # noqa: E402

# (M, N, K, act, residual): one 32-image chunk of CvT-21
SHAPES = [
    (294912, 64, 152, 0, 0), (294912, 64, 64, 0, 0), (73728, 64, 64, 0, 0), (294912, 64, 64, 0, 1), (294912, 256, 64, 1, 0),
    (294912, 64, 256, 0, 1),
    (73728, 192, 576, 0, 0), (73728, 192, 192, 0, 0), (18432, 192, 192, 0, 0), (73728, 192, 192, 0, 1), (73728, 768, 192, 1, 0),
    (73728, 192, 768, 0, 1),
    (18432, 384, 1728, 0, 0), (18464, 384, 384, 0, 0), (4640, 384, 384, 0, 0), (18464, 384, 384, 0, 1), (18464, 1536, 384, 1, 0),
    (18464, 384, 1536, 0, 1), (18432, 768, 384, 0, 0),
]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for M, N, K, act, res in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        R = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
        for _ in range(int(os.environ.get('WARM', '3'))):
            gemm_hook("tcgen05", A, W, bias, act, R, False)
        ts = []
        reps = int(os.environ.get('REPS', '20'))      # back to back: the launch latency of the host is hidden, A (just written / read) is L2-warm as in the encoder
        for _ in range(int(os.environ.get('ITERS', '3'))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _r in range(reps):
                gemm_hook("tcgen05", A, W, bias, act, R, False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        us = sorted(ts)[len(ts) // 2]
        tot += us
        print(f"[{M:6d} x {N:4d} x {K:4d}] act={act} res={res}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TF/s")
    print(f"sum {tot:.0f} us")


if __name__ == "__main__":
    main()
