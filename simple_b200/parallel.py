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
    """Callable for `Trainer.grad_sync`: gradients live in one flat fp32 arena that is all-reduced (mean) and handed to
    `Adafactor.step(grads=...)` as {param: view into the arena} -- no copy back.

    Overlap (packed-gradient mode, see ops.PackedWeight): `prepare()` points every conv layer's wgrad accumulator
    (`pw.grad_slot`) INTO the arena, laid out in the order the backward pass finishes them (decoder first), split into
    buckets.  When the last layer of a bucket reports its gradient (`pw.on_ready`), the bucket's all-reduce is launched
    on a side stream and runs over NVLink while the rest of the backward pass keeps the SMs busy; only the last bucket
    (down1, down0, embed, gate + the 2.3 M non-conv parameters: 4 M of 68 M values) is exposed.  Works under CUDA-graph capture (the side stream
    is forked from / joined to the capturing stream with events).

    Parameters whose gradient is None on this step (gate.* on the first step of a window) are None
    on every rank alike (the condition depends on the step index only), so all ranks see the same layout.
    """

    # backward completion order of the conv layers (autograd runs the decoder first, the gate conv last)
    ORDER = ['logits', 'up4', 'up3', 'up2', 'up1', 'up0', 'mid1', 'mid0', 's_conv2', 's_conv1', 's_embed',
             'down4', 'down3', 'down2', 'down1', 'down0', 'embed', 'gate']

    # layers after which a bucket is sent (backward order): 25 M | 15 M | 10 M | 9 M | 5 M values, 4 M left for the end
    BOUNDS_AFTER = ('up0', 's_conv2', 'down4', 'down3', 'down2')

    def __init__(self, model, group=None, bounds_after=None):
        self.params = [p for p in model.parameters()]
        self.group = group
        self.flat = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        env = os.environ.get('SB_DP_BOUNDS')
        self.bounds_after = tuple(env.split(',')) if env else tuple(bounds_after or self.BOUNDS_AFTER)
        self._layout = None      # overlap mode: dict(name -> (offset, numel)), bucket boundaries, tail offset
        self._comm_stream = None
        self.last_scale = 1.0  # the gradients returned by the last call are (mean gradient) / last_scale
        # SMs the compute kernels leave to the all-reduce CTAs while a bucket is in flight (sb_set_sm_reserve).  Measured at
        # 2 GPUs (tools/dp_matrix.sh): 0 -> 7.19 ms, 4 -> 7.23, 8 -> 7.29, 16 -> 7.31 with NCCL's default CTA count, and no
        # better than 7.18 with any NCCL_MAX_CTAS: the overlap cost is memory-system contention, not displaced CTAs -> off
        self.sm_reserve = int(os.environ.get('SB_DP_SM_RESERVE', '0'))

    # ---- overlap mode ------------------------------------------------------------------------------------------
    @classmethod
    def bucket_layout(cls, sizes, bounds_after):
        """Arena layout of the conv gradients: `sizes` {layer name: packed gradient length} -> slots in backward
        completion order (`ORDER`, unknown layers last) and the buckets.  A bucket ends after each layer named in
        `bounds_after`: few large buckets while plenty of backward is left to hide them behind, then small ones, so that
        what is still in flight when the backward pass ends (the layers after the final boundary plus the non-conv tail:
        ~4 M of 68 M values) is as short as it can be."""
        names = [n for n in cls.ORDER if n in sizes] + [n for n in sizes if n not in cls.ORDER]
        off, slots = 0, {}
        for n in names:
            slots[n] = (off, sizes[n])
            off += sizes[n]
        bounds, members, cur = [], [], set()
        for n in names:
            cur.add(n)
            if n in bounds_after:
                bounds.append(slots[n][0] + slots[n][1])
                members.append(cur)
                cur = set()
        return dict(names=names, slots=slots, conv_total=off, bounds=bounds, members=members)

    def prepare(self, model, packed):
        """Call after the forward pass and before backward (packed-gradient mode only)."""
        if not packed or self.world < 2 or not torch.cuda.is_available():
            return
        pws = model._packed
        dev = next(iter(pws.values())).fwd.device
        if self._layout is None:
            self._layout = self.bucket_layout({n: pw.fwd.numel() for n, pw in pws.items()}, self.bounds_after)
            conv_params = {id(p) for p in packed}
            tail = sum(p.numel() for p in self.params if id(p) not in conv_params)
            self.flat = self._alloc_arena(self._layout['conv_total'] + tail, dev)
            self._comm_stream = torch.cuda.Stream(device=dev)
        L = self._layout
        if self.sm_reserve:  # a step that raised between the first bucket and the end must not leave it set
            from . import _lib
            _lib.lib().sb_set_sm_reserve(0)
        self.flat[:L['conv_total']].zero_()  # wgrad accumulates with red.add
        self._launched = 0
        self._ready = set()
        for n in L['names']:
            o, k = L['slots'][n]
            pw = pws[n]
            pw.grad_slot = self.flat[o:o + k].view(pw.fwd.shape)
            pw.on_ready = (lambda name=n: self._on_ready(name))

    def _alloc_arena(self, n, dev):
        """The gradient arena comes from NCCL's own allocator and is registered with the communicator (user-buffer
        registration): the NVLS all-reduce then reads and writes it in place instead of staging through its own buffers.
        Measured at 8 GPUs: 7.55 -> 7.32 ms per step.  SB_DP_NCCL_POOL=0 falls back to a plain allocation."""
        if os.environ.get('SB_DP_NCCL_POOL', '1') == '1' and dist.get_backend(self.group) == 'nccl':
            try:
                dist.all_reduce(torch.zeros(1, device=dev), group=self.group)  # the communicator must exist first
                backend = (self.group or dist.group.WORLD)._get_backend(dev)
                pool = torch.cuda.MemPool(backend.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    flat = torch.zeros(n, device=dev, dtype=torch.float32)
                backend.register_mem_pool(pool)
                self._pool = pool
                return flat
            except (RuntimeError, AttributeError) as e:  # NCCL built without ncclMemAlloc / torch without MemPool
                import warnings
                warnings.warn(f'FlatGradSync: NCCL-registered gradient arena unavailable ({e}); using a plain allocation')
        return torch.zeros(n, device=dev, dtype=torch.float32)

    def _all_reduce(self, buf):
        # a plain SUM: that is what NCCL can hand to the NVSwitch (NVLS in-switch reduction; AVG is a pre-multiplied sum
        # and falls back to the ring).  The 1 / world is folded into the clip coefficient (`last_scale`).
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)

    def _on_ready(self, name):
        L = self._layout
        self._ready.add(name)
        # a bucket goes out once ALL of its layers have reported (robust to the exact autograd execution order, which is
        # the same on every rank: same graph)
        while self._launched < len(L['bounds']) and L['members'][self._launched] <= self._ready:
            lo = L['bounds'][self._launched - 1] if self._launched else 0
            hi = L['bounds'][self._launched]
            cur = torch.cuda.current_stream()
            if self.sm_reserve:
                from . import _lib
                _lib.lib().sb_set_sm_reserve(self.sm_reserve)
            self._comm_stream.wait_stream(cur)  # this bucket's gradients are complete on the compute stream
            with torch.cuda.stream(self._comm_stream):
                self._all_reduce(self.flat[lo:hi])
            self._launched += 1

    def _finish_overlapped(self, packed, overlap=None):
        L = self._layout
        conv = {id(p): pw for p, pw in packed.items()}
        rest = [p for p in self.params if id(p) not in conv and p.grad is not None]
        views, off = {}, L['conv_total']
        for p in rest:
            views[p] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if rest:
            torch._foreach_copy_(list(views.values()), [p.grad for p in rest])
        lo = L['bounds'][self._launched - 1] if self._launched else 0
        # the not-yet-launched buckets + the non-conv tail: nothing of the backward pass is left to hide them behind, so
        # the caller's `overlap` work (frame feedback, state commit: independent of the gradients) runs meanwhile
        cur = torch.cuda.current_stream()
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            self._all_reduce(self.flat[lo:off])
        if overlap is not None:
            overlap()
        cur.wait_stream(self._comm_stream)
        if self.sm_reserve:
            from . import _lib
            _lib.lib().sb_set_sm_reserve(0)
        for p, pw in packed.items():
            if pw.grad_slot is not None and pw.gpacked is not None:
                views[p] = pw.grad_slot
            pw.on_ready = None
        self.last_scale = 1.0 / self.world
        return views

    # ---- entry point -------------------------------------------------------------------------------------------
    def __call__(self, model, packed=None, overlap=None):
        """`packed`: optional {param: ops.PackedWeight}; those parameters contribute their packed-layout gradient
        (`pw.gpacked`, zero-padded, identical layout on every rank) instead of `p.grad`.  `overlap`: optional callable
        with gradient-independent work of the step, run while the last all-reduce is in flight."""
        if packed and self._layout is not None and all(pw.grad_slot is not None for pw in packed.values()):
            return self._finish_overlapped(packed, overlap)
        if overlap is not None:
            overlap()
        self.last_scale = 1.0
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


def seed_rank_streams(model, base_seed=0):
    """Per-rank random streams (SURVEY 8e caveat 3): the replicas hold identical weights but must draw INDEPENDENT dropout
    masks / latent noise for their shards, like the samples of one global batch do.  Offsets the model's in-kernel
    dropout seed and torch's CUDA generator by the rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = model.logits.weight.device
    seed = (int(base_seed) + 0x9E3779B9 * (rank + 1)) & ((1 << 40) - 1)
    if model._seed_dev is not None and model._seed_dev.device == dev:
        model._seed_dev.fill_(seed)
    else:
        model._seed_dev = torch.full((1,), seed, dtype=torch.int64, device=dev)
    torch.cuda.manual_seed(seed)


def broadcast_parameters(model, src=0):
    """Make every rank hold rank `src`'s weights at start-up (same effect as seeding identically).  Not needed afterwards:
    the reduced gradient is bit-identical on every rank and every reduction of the fused clip + Adafactor kernels is
    order-deterministic, so the replicas never drift (asserted in tests/test_dp_nccl.py)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src)
        packed = getattr(model, '_packed', None)
        if packed:
            for pw in packed.values():
                pw.fresh = False  # the bf16 GEMM operands are re-derived from the synchronised masters
