"""Optional per-launch CUDA-event timing of the hot kernels (used by bench.py's roofline pass; off by default)."""
import torch

RECORDER = None


class Recorder:
    def __init__(self):
        self.items = []  # (kernel, tag, work, unit, ev0, ev1)

    def start(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def stop(self, kernel, tag, work, unit, ev0):
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        self.items.append((kernel, tag, work, unit, ev0, ev1))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kernel, tag, work, unit, e0, e1 in self.items:
            d = out.setdefault(kernel, {'ms': 0.0, 'work': 0.0, 'launches': 0, 'unit': unit, 'by_tag': {}})
            ms = e0.elapsed_time(e1)
            d['ms'] += ms
            d['work'] += work
            d['launches'] += 1
            t = d['by_tag'].setdefault(tag, [0.0, 0.0, 0])
            t[0] += ms
            t[1] += work
            t[2] += 1
        return out
