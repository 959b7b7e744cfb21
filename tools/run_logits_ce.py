"""Runs the fused logits + softmax-CE kernel (and, with `unfused`, the two-kernel path) a few times at B = 64: a target
for ncu / timing.  python tools/run_logits_ce.py [fused|unfused|eval] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_logits_ce import _setup  # noqa: E402
from simple_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'fused'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg, model, hf, target = _setup(B)
spec, pw, m = model._plan['logits'], model._packed['logits'], model.logits
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
with torch.no_grad():
    for it in range(8):
        ev[0].record()
        if mode == 'fused':
            ops.LogitsCE.apply(spec, hf, m.weight, m.bias, pw, target, 0.03)
        elif mode == 'eval':
            ops.logits_argmax(spec, hf, m.bias, pw)
        else:
            lg = ops.tap_conv(spec, hf, m.weight, m.bias, pw)
            ops.PixelCE.apply(lg, target, 0.03, True, True, None)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
print(mode, 'B', B, 'ms per call (incl. small torch ops):', ['%.3f' % t for t in times])
