"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (NVLink 5 / NVSwitch).

World-model training is data parallel: the global batch is sharded, every rank runs the same kernels on its shard,
and ONE all-reduce of the flat gradient (68.34 M fp32 = 273 MB) per optimiser step makes the replicas identical; the
fused clip+Adafactor then reads the reduced gradient straight out of the flat buffer (no copy back).  Simulated
rollouts shard by agent and need no collective (`simulated_env.shard_agents`).  The reference has no distributed
code at all; semantics follow what a global-batch run of the reference would compute: the losses are
means over the batch, so the global gradient is the mean of the per-shard gradients.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world_size, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


class FlatGradSync:
    """Callable for `Trainer.grad_sync`: packs the gradients into one flat fp32 buffer, all-reduces it (mean), and
    returns {param: view into the buffer} for `Adafactor.step(grads=...)`.

    Parameters whose gradient is None on this step (gate.* on the first step of a window) are None
    on every rank alike (the condition depends on the step index only), so all ranks pack the same layout.
    """

    def __init__(self, model, group=None):
        self.params = [p for p in model.parameters()]
        self.group = group
        self.flat = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, model, packed=None):
        """`packed`: optional {param: ops.PackedWeight}; those parameters contribute their packed-layout gradient
        (`pw.gpacked`, zero-padded, identical layout on every rank) instead of `p.grad`."""
        active = []
        for p in self.params:
            pw = packed.get(p) if packed else None
            g = pw.gpacked if pw is not None and pw.gpacked is not None else p.grad
            if g is not None:
                active.append((p, g))
        n = sum(g.numel() for _, g in active)
        if self.flat is None or self.flat.numel() < n:
            self.flat = torch.empty(n + (1 << 20), device=active[0][1].device, dtype=torch.float32)
        views = {}
        off = 0
        for p, g in active:
            views[p] = self.flat[off:off + g.numel()].view(g.shape)
            off += g.numel()
        torch._foreach_copy_(list(views.values()), [g for _, g in active])
        if self.world > 1:
            buf = self.flat[:n]
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.mul_(1.0 / self.world)
        return views


def broadcast_parameters(model, src=0):
    """Make every rank start from rank `src`'s weights (same effect as seeding identically)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src)
