"""Adafactor (drop-in for simple/adafactor.py of the reference, used with all defaults at simple/trainer.py:22).

Same constructor / `zero_grad()` / `step()` surface and the same per-parameter state keys (`step`,
`exp_avg_sq_row`, `exp_avg_sq_col` or `exp_avg_sq`, `RMS`), but the whole update -- including the preceding
`clip_grad_norm_` when `step(max_grad_norm=...)` is used -- runs in three fused multi-tensor CUDA kernels
(csrc/adafactor.cu) with no host synchronisation.  Only the defaults the reference uses are implemented; any other
hyper-parameter raises (there is no slow path).
"""

import numpy as np
import torch

from . import _lib

_DESC = np.dtype([('p', np.uint64), ('g', np.uint64), ('row', np.uint64), ('col', np.uint64), ('acc', np.uint64),
                  ('n', np.int64), ('rows', np.int32), ('cols', np.int32), ('stepf', np.uint64)])
assert _DESC.itemsize == 64
_PACK = np.dtype([('gpacked', np.uint64), ('fwd', np.uint64), ('iperm', np.uint64), ('operm', np.uint64),
                  ('O', np.int32), ('I', np.int32), ('s', np.int32), ('Ipad', np.int32), ('Opad', np.int32),
                  ('pad_', np.int32, (3,))])
assert _PACK.itemsize == 64


class Adafactor(torch.optim.Optimizer):

    def __init__(self, params, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                 weight_decay=0.0, scale_parameter=True, relative_step=True, warmup_init=False):
        if lr is not None and relative_step:
            raise ValueError('Cannot combine manual lr and relative_step options')
        if warmup_init and not relative_step:
            raise ValueError('warmup_init requires relative_step=True')
        if (lr is not None or tuple(eps) != (1e-30, 1e-3) or clip_threshold != 1.0 or decay_rate != -0.8
                or beta1 is not None or weight_decay != 0.0 or not scale_parameter or not relative_step or warmup_init):
            raise ValueError('simple_b200.Adafactor implements the defaults used by the reference only')
        defaults = dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate, beta1=beta1,
                        weight_decay=weight_decay, scale_parameter=scale_parameter, relative_step=relative_step,
                        warmup_init=warmup_init)
        super().__init__(params, defaults)
        self._cache = {}
        self.grad_norm_sq = None  # device scalar: squared global gradient norm of the last step (before clipping)
        self._grad_scale = 1.0

    @property
    def grad_scale(self):
        """The gradients handed to `step` are (true gradient) / grad_scale.  Data-parallel training sets 1 / world: the
        shards' gradients are SUMMED (a plain sum is what the NVSwitch can reduce), the clip coefficient is computed from
        the true norm (`sb_grad_sumsq` scales the sum of squares), and nothing else needs to know -- Adafactor's update
        g / sqrt(EMA g^2) is invariant to a constant gradient scale (bit-exactly so for a power of two; eps1 = 1e-30
        aside).  The second-moment states live in the same units; changing the scale re-scales them."""
        return self._grad_scale

    @grad_scale.setter
    def grad_scale(self, s):
        s = float(s)
        if s <= 0:
            raise ValueError('grad_scale must be positive')
        if s != self._grad_scale:
            f = (self._grad_scale / s) ** 2
            for st in self.state.values():
                for k in ('exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq'):
                    if k in st:
                        st[k].mul_(f)
            self._grad_scale = s

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = 0
            if p.dim() >= 2:
                st['exp_avg_sq_row'] = torch.zeros(p.shape[:-1], device=p.device, dtype=torch.float32)
                st['exp_avg_sq_col'] = torch.zeros(p.shape[:-2] + p.shape[-1:], device=p.device, dtype=torch.float32)
            else:
                st['exp_avg_sq'] = torch.zeros_like(p, dtype=torch.float32)
            st['RMS'] = 0
            st['step_dev'] = torch.zeros(1, device=p.device, dtype=torch.float32)  # device mirror of `step`
        return st

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None, grads=None, packed=None):
        """One optimisation step.  `max_grad_norm` fuses torch.nn.utils.clip_grad_norm_(params, max_grad_norm) in
        front (simple/trainer.py:148); `grads` optionally maps parameters to externally held gradient tensors
        (e.g. views of an all-reduced flat buffer) instead of `p.grad`; `packed` optionally maps 4-D parameters to
        their `ops.PackedWeight`: their gradient is then read in the packed GEMM-operand layout (`pw.gpacked`, or
        `grads[p]` holding a tensor of that layout) and the bf16 operands are refreshed by the update itself."""
        loss = closure() if closure is not None else None
        active = []
        for group in self.param_groups:
            for p in group['params']:
                pw = packed.get(p) if packed is not None else None
                g = grads.get(p) if grads is not None else None
                if g is None:
                    g = pw.gpacked if pw is not None and pw.gpacked is not None else (p.grad if grads is None else None)
                if pw is not None and (g is None or g.dim() != 2):
                    pw = None  # reference-layout gradient: plain path
                if g is None:
                    continue  # adafactor.py:125-126 -- e.g. gate.* on the first step of a rollout window
                if not p.is_cuda:
                    raise RuntimeError('simple_b200.Adafactor runs on CUDA only; there is no CPU fallback')
                if p.dtype != torch.float32 or g.dtype != torch.float32 or p.dim() not in (1, 2, 4):
                    raise RuntimeError('unsupported parameter for the fused Adafactor')
                if not g.is_contiguous():
                    g = g.contiguous()
                    if grads is None and pw is None:
                        p.grad = g
                assert p.is_contiguous()
                active.append((p, g, pw))
        if not active:
            return loss
        dev = active[0][0].device
        n = len(active)
        key = tuple((id(p), pw is not None) for p, _, pw in active)
        rec = self._cache.get(key)
        if rec is None:
            # per parameter-set record: pinned host descriptors (CUDA-graph replays re-read them), device copies,
            # accumulators and block maps.  Two sets exist in practice: with and without gate.* (first step of a window).
            # host layout: [n x sb_tensor_desc (64 B)] [n x sb_pack_desc (64 B)] [n x int64 gradient length]
            nbytes = n * (64 + 64 + 8)
            host = torch.zeros(nbytes, dtype=torch.uint8)
            hn = host.numpy()
            rec = dict(host=host, np=hn[:n * 64].view(_DESC), packs=hn[n * 64:n * 128].view(_PACK),
                       glen=hn[n * 128:].view(np.int64), dev=torch.zeros(nbytes, dtype=torch.uint8, device=dev),
                       acc=torch.zeros(n, 2, device=dev, dtype=torch.float32),
                       total=torch.zeros(1, device=dev, dtype=torch.float32), maps=self._build_maps(active, dev),
                       # scratch of the order-deterministic reductions (ticket counters zeroed once; see the header)
                       sumsq_scratch=torch.zeros(1 + 148 * 8, device=dev, dtype=torch.float32),
                       params=[p for p, _, _ in active], pin=torch.zeros(nbytes, dtype=torch.uint8).pin_memory(),
                       any_packed=any(pw is not None for _, _, pw in active))
            self._cache[key] = rec
        descs, packs, glen = rec['np'], rec['packs'], rec['glen']
        acc, total = rec['acc'], rec['total']
        for i, (p, g, pw) in enumerate(active):
            st = self._init_state(p)
            st['step'] += 1
            d = descs[i]
            d['p'], d['g'] = p.data_ptr(), g.data_ptr()
            if p.dim() >= 2:
                d['row'], d['col'] = st['exp_avg_sq_row'].data_ptr(), st['exp_avg_sq_col'].data_ptr()
            else:
                d['row'] = st['exp_avg_sq'].data_ptr()
            d['acc'] = acc.data_ptr() + 8 * i
            d['n'] = p.numel()
            d['rows'], d['cols'] = (p.shape[0], p.shape[1]) if p.dim() == 2 else (0, 0)
            d['stepf'] = st['step_dev'].data_ptr()
            glen[i] = g.numel()
            if pw is not None:
                assert tuple(g.shape) == tuple(pw.fwd.shape)
                k = packs[i]
                k['gpacked'], k['fwd'] = g.data_ptr(), pw.fwd.data_ptr()
                k['iperm'] = pw.iperm.data_ptr() if pw.iperm is not None else 0
                k['operm'] = pw.operm.data_ptr() if pw.operm is not None else 0
                k['O'], k['I'], k['s'], k['Ipad'], k['Opad'] = pw.O, pw.I, pw.s, pw.Ipad, pw.Opad
            st['RMS'] = _LazyRMS(acc, i, p.numel())
        if 'conv_scratch' not in rec:
            nb = max([bm.shape[0] for bm in rec['maps']['conv'].values()] + [1])
            rec['conv_scratch'] = torch.zeros(n + 2 * nb, device=dev, dtype=torch.float32)
        self._launch(rec, n, max_grad_norm)
        from .ops import refresh_transposed_all
        refresh_transposed_all([pw for _, _, pw in active], rec)  # `fwd` was rewritten: re-derive the per-tap transposes
        self.last_key = key
        return loss

    def advance_for_replay(self, key):
        """Host-side bookkeeping for one replay of a CUDA graph that captured `step()` on parameter set `key`: the
        device step counters advance inside the graph; this keeps the reference-visible `state[p]['step']` in sync."""
        for p in self._cache[key]['params']:
            self.state[p]['step'] += 1

    def _launch(self, rec, n, max_grad_norm):
        maps, acc, total, dev_buf = rec['maps'], rec['acc'], rec['total'], rec['dev']
        # Descriptors hold pointers only.  Pageable source => the copy is synchronous w.r.t. the host buffer, so the next
        # step may rewrite it; under CUDA-graph capture the pointers are static and the device copy is already valid.
        if torch.cuda.is_current_stream_capturing():
            # gradients allocated during capture live at graph-private addresses: upload them through a pinned staging
            # buffer whose content never changes afterwards (every replay re-issues the same small copy)
            rec['pin'].copy_(rec['host'])
            dev_buf.copy_(rec['pin'], non_blocking=True)
        else:
            dev_buf.copy_(rec['host'])
        acc.zero_()
        total.zero_()
        L = _lib.lib()
        s = _lib.stream_ptr()
        descs_ptr = dev_buf.data_ptr()
        packs_ptr = descs_ptr + n * 64
        glen_ptr = descs_ptr + n * 128
        _lib.check(L.sb_adafactor_tick(descs_ptr, n, s), 'sb_adafactor_tick')
        mn = float(max_grad_norm) if max_grad_norm is not None else 0.0
        _lib.check(L.sb_grad_sumsq(descs_ptr, glen_ptr, maps['sumsq'].data_ptr(), maps['sumsq'].shape[0],
                                   total.data_ptr(), rec['sumsq_scratch'].data_ptr(), float(self.grad_scale) ** 2, s),
                   'sb_grad_sumsq')
        cs = rec['conv_scratch'].data_ptr()
        for (kh, kw, is_packed), bm in maps['conv'].items():
            for pas in (1, 2):
                if is_packed:
                    _lib.check(L.sb_adafactor_conv_packed(descs_ptr, packs_ptr, bm.data_ptr(), bm.shape[0], kh, kw,
                                                          total.data_ptr(), mn, pas, cs, n, s),
                               'sb_adafactor_conv_packed')
                else:
                    _lib.check(L.sb_adafactor_conv(descs_ptr, bm.data_ptr(), bm.shape[0], kh, kw, total.data_ptr(), mn,
                                                   pas, cs, n, s), 'sb_adafactor_conv')
        if maps['big2d'] is not None:
            idx, mc, mr = maps['big2d']
            _lib.check(L.sb_adafactor_big2d(descs_ptr, idx.data_ptr(), idx.numel(), mc, mr, total.data_ptr(), mn, s),
                       'sb_adafactor_big2d')
        if maps['small'] is not None:
            idx, rc = maps['small']
            _lib.check(L.sb_adafactor_small(descs_ptr, idx.data_ptr(), idx.numel(), rc, total.data_ptr(), mn, s),
                       'sb_adafactor_small')
        self.grad_norm_sq = total

    @staticmethod
    def _build_maps(active, dev):
        """Launch geometry of one parameter set.  Every map refers to descriptors by their index in `active`."""
        sumsq = []
        conv = {}
        small = []
        big = []
        rc = 1
        for i, (p, g, pw) in enumerate(active):
            nchunks = -(-g.numel() // 4096)
            sumsq += [(i, c) for c in range(nchunks)]
            if p.dim() == 4:
                conv.setdefault((p.shape[2], p.shape[3], pw is not None), []).append(i)
            elif p.dim() == 2 and p.numel() >= 49152 and p.shape[0] >= 8:
                big.append(i)  # one 8-CTA cluster per tensor instead of a single CTA
            else:
                small.append(i)
                if p.dim() == 2:
                    rc = max(rc, p.shape[0] + p.shape[1])
        maps = {'sumsq': torch.tensor(sumsq, dtype=torch.int32, device=dev), 'conv': {}, 'small': None, 'big2d': None}
        if big:
            maps['big2d'] = (torch.tensor(big, dtype=torch.int32, device=dev),
                             max(active[i][0].shape[1] for i in big), max(active[i][0].shape[0] for i in big))
        for k, idxs in conv.items():
            bm = []
            for i in idxs:
                npatch = active[i][0].numel() // (k[0] * k[1])
                bm += [(i, b * 128) for b in range(-(-npatch // 128))]
            maps['conv'][k] = torch.tensor(bm, dtype=torch.int32, device=dev)
        if small:
            maps['small'] = (torch.tensor(small, dtype=torch.int32, device=dev), rc)
        return maps


class _LazyRMS:
    """state['RMS'] of the reference is a 0-d tensor; here it is derived on demand from the device accumulators so
    that the optimiser never synchronises the host."""

    def __init__(self, acc, i, n):
        self.acc, self.i, self.n = acc, i, n

    def tensor(self):
        return torch.sqrt(self.acc[self.i, 1] / self.n)

    def __float__(self):
        return float(self.tensor())
