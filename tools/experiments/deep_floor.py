"""Latency floor of the split-K tap GEMM on a deep-layer shape (B x 3 x 2 pixels, K = 16 taps x 768, N = 768):
time per launch against the number of K steps per CTA.  python tools/experiments/deep_floor.py [B]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from simple_b200.gemm import tap_gemm  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
torch.manual_seed(0)
C, N, T = 768, 768, 16
a = torch.randn(B, 3, 2, C, device=dev).bfloat16()
w = (torch.randn(N, T * C, device=dev) * 0.02).bfloat16()
taps = [(0, 0)] * T  # same pixel for every tap: the traffic and the K loop of a 4x4 conv, no geometry
out = torch.zeros(B, 3, 2, N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for splits in (1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 192):
    ts = []
    for it in range(12):
        flush.zero_()  # weights out of L2, like in the real step
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tap_gemm(a, w, taps, 3, 2, N=N, out=out, splits=splits, accumulate=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    ksteps = T * (C // 64) / splits
    print(f'splits {splits:4d}  K steps per CTA {ksteps:6.1f}  CTAs {min(148, 3 * splits):4d}  median {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f} us')
