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

    def add_work(self, kernel, tag, work, unit, timed_item):
        """Credit `work` of another kind (e.g. the GEMM flops of a fused GEMM + loss launch) to the events of an item
        that was already recorded."""
        self.items.append((kernel, tag, work, unit, timed_item[4], timed_item[5]))

    def summary(self):
        """Per kernel: total ms / work / launches, and per tag [ms, work, launches].  A tag's time is the MEDIAN of its
        samples times their count: one preempted or cold launch (seen: a single 1.8 ms outlier among 80 us launches)
        must not move a roofline figure."""
        torch.cuda.synchronize()
        samples = {}
        for kernel, tag, work, unit, e0, e1 in self.items:
            samples.setdefault((kernel, tag, unit), []).append((e0.elapsed_time(e1), work))
        out = {}
        for (kernel, tag, unit), lst in samples.items():
            d = out.setdefault(kernel, {'ms': 0.0, 'work': 0.0, 'launches': 0, 'unit': unit, 'by_tag': {}})
            times = sorted(ms for ms, _ in lst)
            med = times[len(times) // 2] if len(times) % 2 else 0.5 * (times[len(times) // 2 - 1] + times[len(times) // 2])
            ms, work = med * len(lst), sum(w for _, w in lst)
            d['ms'] += ms
            d['work'] += work
            d['launches'] += len(lst)
            d['by_tag'][tag] = [ms, work, len(lst)]
        return out
