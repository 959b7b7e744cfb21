"""A/B timing of the pixel-CE kernel variants on B=64 synthetic logits (CUDA events, L2 flushed by size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_b200 import _lib, ops
dev = 'cuda:0'
B, H, W = 64, 105, 80
torch.manual_seed(0)
base = (torch.randn(B, H, W, 768, device=dev) * 3).bfloat16()
tgt = torch.randint(0, 256, (B, 3, H, W), device=dev, dtype=torch.uint8)
L = _lib.lib()
for mode in (0, 2, 1, 0, 2):
    L.sb_set_ce_tma(mode)
    times = []
    for it in range(8):
        lg = base.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, am = ops.PixelCE.apply(lg, tgt, 0.03, True, True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = sorted(times[2:])[len(times[2:]) // 2]
    print(f'mode {mode}: {t * 1e3:.1f} us  {1654732800 / t / 1e6:.0f} GB/s  loss {float(loss):.4f}')
L.sb_set_ce_tma(0)
