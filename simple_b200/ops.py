"""Host-side operators of the world-model hot path: thin autograd wrappers over the C-ABI CUDA kernels.

Every operator here launches hand-written sm_100a kernels through `simple_b200._lib` and raises if the library is
missing (no CPU / ATen fallback for the kernels).  torch supplies allocation, streams and the autograd tape.
"""
import ctypes
import os

import torch

from . import _lib, profiling
from .gemm import tap_gemm, tap_wgrad


# ---------------------------------------------------------------------------------------------------------------
# layouts
# ---------------------------------------------------------------------------------------------------------------
class Lay:
    """Physical layout of a logical [B,H,W,C] bf16 activation: plain NHWC (s=1) or padded space-to-depth."""

    def __init__(self, H, W, C, s=1, pad=0, H2=None, W2=None):
        self.H, self.W, self.C, self.s, self.pad = H, W, C, s, pad
        self.kind = 0 if s == 1 else 1
        self.H2 = H2 if H2 is not None else H
        self.W2 = W2 if W2 is not None else W

    def shape(self, B):
        if self.kind == 0:
            return (B, self.H, self.W, self.C)
        return (B, self.H2, self.W2, self.s * self.s * self.C)

    def view(self, t):
        v = _lib.SbView()
        v.p = t.data_ptr() if t is not None else None
        v.kind, v.s, v.pad, v.H2, v.W2 = self.kind, self.s, self.pad, self.H2, self.W2
        return v

    def to_nchw(self, t):
        """Debug/test helper: physical tensor -> logical [B,C,H,W] fp32."""
        B = t.shape[0]
        if self.kind == 0:
            return t.float().permute(0, 3, 1, 2)
        s = self.s
        x = t.float().reshape(B, self.H2, self.W2, s, s, self.C).permute(0, 5, 1, 3, 2, 4)
        x = x.reshape(B, self.C, self.H2 * s, self.W2 * s)
        return x[:, :, self.pad:self.pad + self.H, self.pad:self.pad + self.W]


def plain(H, W, C):
    return Lay(H, W, C)


def s2d_in(H, W, C, k, s, pad):
    """Layout consumed by a k x k / stride s / padding pad convolution expressed as (k/s)^2 stride-1 taps."""
    Ho = (H + 2 * pad - k) // s + 1
    Wo = (W + 2 * pad - k) // s + 1
    return Lay(H, W, C, s=s, pad=pad, H2=Ho + k // s - 1, W2=Wo + k // s - 1)


def s2d_out(Hin, Win, Hout, Wout, C):
    """Layout produced by a 4x4 / stride 2 / padding 1 transposed convolution run as 4 taps: N = (phase, C)."""
    return Lay(Hout, Wout, C, s=2, pad=1, H2=Hin + 1, W2=Win + 1)


def _ptr(t):
    return t.data_ptr() if t is not None else None


class ZeroArena:
    """All zero-initialised scratch of one training step (InstanceNorm sums, split-K / weight-gradient accumulators,
    bias-gradient and loss accumulators: ~80 buffers, 280 MB at B = 64) carved from ONE buffer that a single memset
    clears at the start of the step, instead of one fill kernel per buffer.  `begin()` .. `end()` brackets a step; the
    first bracketed step measures the demand (requests fall back to torch.zeros), later steps are served from the arena.
    Tensors handed out stay valid until the next `begin()`."""

    def __init__(self):
        self.buf, self.off, self.need, self.active = None, 0, 0, False

    def begin(self, device):
        want = self.need + (self.need >> 3)
        if want > 0 and (self.buf is None or self.buf.numel() < want or self.buf.device != torch.device(device)):
            self.buf = torch.empty((want + 15) & ~15, dtype=torch.uint8, device=device)
        if self.buf is not None:
            self.buf.view(torch.float32).zero_()  # 16-byte stores (a uint8 fill writes 4 bytes per thread: 90 -> ~50 us)
        self.off, self.active, self._demand = 0, True, 0

    def end(self):
        self.active = False
        self.need = max(self.need, getattr(self, '_demand', 0))

    def zeros(self, shape, dtype, device):
        n = 1
        for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
            n *= int(d)
        nbytes = (n * torch.empty(0, dtype=dtype).element_size() + 255) & ~255
        if self.active:
            self._demand += nbytes
            if self.buf is not None and self.off + nbytes <= self.buf.numel() and self.buf.device == torch.device(device):
                t = self.buf[self.off:self.off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(shape)
                self.off += nbytes
                return t
        return torch.zeros(shape, dtype=dtype, device=device)


ARENA = ZeroArena()


def zeros(shape, device, dtype=torch.float32):
    return ARENA.zeros(shape, dtype, device)


class use_arena:
    """`with use_arena(a):` serves `zeros()` from another arena (the simulated env keeps its own: a buffer baked into its
    captured step must not be re-allocated when the trainer's demand grows)."""

    def __init__(self, arena):
        self.arena = arena

    def __enter__(self):
        global ARENA
        self.prev, ARENA = ARENA, self.arena
        return self.arena

    def __exit__(self, *exc):
        global ARENA
        ARENA = self.prev
        return False


def dense(a, b, a_trans=False, b_trans=False, bias=None, out=None):
    """fp32 `op(a) @ op(b) (+ bias)`, or `out += ...` when `out` is given (sb_dense: the skinny GEMMs of the latent /
    heads block).  a: [M,K] ([K,M] with a_trans); b: [K,N] ([N,K] with b_trans, i.e. an nn.Linear weight); unit inner
    strides, any row pitch (views into the flat parameter banks).  Small outputs are split along K across the chip and
    summed with red.add into zero-initialised memory (the step's zero arena)."""
    for t in (a, b):
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError('sb_dense: operands must be 2-D fp32 with a unit inner stride')
    K, M = a.shape if a_trans else (a.shape[1], a.shape[0])
    N = b.shape[0] if b_trans else b.shape[1]
    if (b.shape[1] if b_trans else b.shape[0]) != K:
        raise ValueError(f'sb_dense: inner dimensions differ ({tuple(a.shape)} x {tuple(b.shape)})')
    L = _lib.lib()
    if out is not None:
        if out.dtype != torch.float32 or out.shape != (M, N) or out.stride(1) != 1:
            raise ValueError('sb_dense: bad accumulator')
        acc = 1
    elif L.sb_dense_splits(M, N, K) > 1:
        out, acc = zeros((M, N), a.device), 1
    else:
        out, acc = torch.empty(M, N, device=a.device), 0
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError('sb_dense: bad bias')
    _lib.check(L.sb_dense(a.data_ptr(), b.data_ptr(), out.data_ptr(), bias.data_ptr() if bias is not None else None,
                          M, N, K, a.stride(0), b.stride(0), out.stride(0), int(a_trans), int(b_trans), acc,
                          _lib.stream_ptr()), 'sb_dense')
    return out


# ---------------------------------------------------------------------------------------------------------------
# raw kernel wrappers
# ---------------------------------------------------------------------------------------------------------------
class PrepSpec:
    """Static description of one fused elementwise stage (see sb_prep in include/simple_b200.h)."""

    def __init__(self, H, W, C, lay_in, lay_out, relu_in=False, norm=False, Ta=None, want_out1=False, p=0.0,
                 stream_id=0, Tb=None, inject=False, eps=1e-6):
        self.H, self.W, self.C = H, W, C
        self.lay_in, self.lay_out = lay_in, lay_out
        self.relu_in, self.norm, self.Ta, self.want_out1 = relu_in, norm, Ta, want_out1
        self.p, self.stream_id, self.Tb, self.inject, self.eps = p, stream_id, Tb, inject, eps


def _prep_args(spec, B, main, add, sums, gamma, beta, sc, sh, out1, out2, seed, keep):
    a = _lib.PrepArgs()
    a.B, a.H, a.W, a.C = B, spec.H, spec.W, spec.C
    a.inp = spec.lay_in.view(main)
    a.add = _ptr(add)
    a.relu_in = int(spec.relu_in)
    a.in_f32 = int(main is not None and main.dtype == torch.float32)
    a.add_f32 = int(add is not None and add.dtype == torch.float32)
    a.sums = _ptr(sums)
    a.eps = spec.eps
    a.gamma, a.beta = _ptr(gamma), _ptr(beta)
    a.Ta = _ptr(spec.Ta)
    a.out1 = _ptr(out1)
    a.p = float(spec.p)
    if isinstance(seed, torch.Tensor):  # device-resident seed (int64 tensor): survives CUDA-graph replay
        a.seed = 0
        a.seed_ptr = seed.data_ptr()
    else:
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    a.stream_id = spec.stream_id
    a.keep = _ptr(keep)
    a.Tb = _ptr(spec.Tb)
    a.sc, a.sh = _ptr(sc), _ptr(sh)
    a.out2 = spec.lay_out.view(out2)
    # tiny planes (the 6x5 / 3x2 bottleneck InstanceNorms): two-pass (mean, variance) statistics, see sb_prep_args
    a.moments = int(bool(spec.norm) and spec.H * spec.W <= 64)
    return a


# Batch chunking (experiment, OFF by default): a stage's two passes (statistics -> apply, backward reduce -> backward
# apply) read the same tensors; run chunk by chunk, the second pass finds its inputs in the 126 MB L2 instead of HBM.
# Measured on B200 at B=64: 8.03 ms/step unchunked, 8.35 ms at 40 MB chunks, 8.89 ms at 24 MB -- the smaller grids lose
# more (tail effects, fewer bytes in flight) than the L2 hits win.  SB_PREP_CHUNK_MB > 0 re-enables it.
CHUNK_BYTES = int(float(os.environ.get('SB_PREP_CHUNK_MB', '0')) * (1 << 20))


def _chunks(B, bytes_per_sample):
    if CHUNK_BYTES <= 0 or B * bytes_per_sample <= 2 * CHUNK_BYTES:
        return [(0, 0)]  # one launch over the whole batch
    nb = max(1, CHUNK_BYTES // max(1, bytes_per_sample))
    return [(b0, min(nb, B - b0)) for b0 in range(0, B, nb)]


def _sample_bytes(*tensors):
    return sum(t.numel() * t.element_size() // t.shape[0] for t in tensors if t is not None)


def stats_raw(spec, main, add=None):
    """Per-(b,c) raw sums [B,C,2] (sum, sum of squares) of relu?(main)+add over the logical H x W pixels."""
    B = main.shape[0]
    sums = zeros((B, spec.C, 2), main.device)
    a = _prep_args(spec, B, main, add, None, None, None, None, None, None, None, 0, None)
    _lib.check(_lib.lib().sb_stats(ctypes.byref(a), sums.data_ptr(), _lib.stream_ptr()), 'sb_stats')
    return sums


class FusedPrep(torch.autograd.Function):
    """z = relu?(main)+add; x = IN(z)*gamma+beta + Ta; out1 = x; out2 = (dropout(x)+Tb)*sc + sh, with layout change."""

    @staticmethod
    def forward(ctx, spec, main, add, gamma, beta, sc, sh, seed, keep, dbias_dst=None):
        B = main.shape[0]
        dev = main.device
        L, st = _lib.lib(), _lib.stream_ptr()
        gamma_f = gamma.detach().float().contiguous() if gamma is not None else None
        beta_f = beta.detach().float().contiguous() if beta is not None else None
        sc_f = sc.detach().float().contiguous() if sc is not None else None
        sh_f = sh.detach().float().contiguous() if sh is not None else None
        out1 = torch.empty((B, spec.H, spec.W, spec.C), device=dev, dtype=torch.bfloat16) if spec.want_out1 else None
        out2 = torch.empty(spec.lay_out.shape(B), device=dev, dtype=torch.bfloat16)
        sums = zeros((B, spec.C, 2), dev) if spec.norm else None
        a = _prep_args(spec, B, main, add, sums, gamma_f, beta_f, sc_f, sh_f, out1, out2, seed, keep)
        a_st = _prep_args(spec, B, main, add, None, None, None, None, None, None, None, 0, None) if spec.norm else None
        # (without InstanceNorm there is a single pass: nothing to keep in the L2 between passes)
        for b0, nb in (_chunks(B, _sample_bytes(main, add)) if spec.norm else [(0, 0)]):
            if spec.norm:
                a_st.b_base, a_st.b_count = b0, nb
                _lib.check(L.sb_stats(ctypes.byref(a_st), sums.data_ptr(), st), 'sb_stats')
            a.b_base, a.b_count = b0, nb
            _lib.check(L.sb_prep(ctypes.byref(a), st), 'sb_prep')
        ctx.spec, ctx.seed, ctx.dbias_dst = spec, seed, dbias_dst
        ctx.save_for_backward(main, add, sums, gamma_f, beta_f, sc_f, sh_f, keep)
        ctx.has = (add is not None, gamma is not None, sc is not None)
        if out1 is None:
            out1 = torch.empty(0, device=dev, dtype=torch.bfloat16)
            ctx.mark_non_differentiable(out1)
        return out1, out2

    @staticmethod
    def backward(ctx, g1, g2):
        spec = ctx.spec
        main, add, sums, gamma_f, beta_f, sc_f, sh_f, keep = ctx.saved_tensors
        B = main.shape[0]
        dev = main.device
        has_add, has_norm, has_inj = ctx.has
        if g2 is None:
            g2 = torch.zeros(spec.lay_out.shape(B), device=dev, dtype=torch.bfloat16)
        g2 = g2.contiguous()
        if not spec.want_out1:
            g1 = None
        if g1 is not None:
            g1 = g1.contiguous()
        a = _prep_args(spec, B, main, add, sums, gamma_f, beta_f, sc_f, sh_f, None, None, ctx.seed, keep)
        a.g2 = spec.lay_out.view(g2)
        a.g1 = _ptr(g1)
        L, st = _lib.lib(), _lib.stream_ptr()
        bsums = dgb = None
        two_pass = has_norm or has_inj
        if two_pass:
            # one zero-filled buffer: per-(b,c) sums [B,C,4] + their batch sums for the affine gradients [C,2]
            buf = zeros((B * spec.C * 4 + spec.C * 2,), dev)
            bsums = buf[:B * spec.C * 4].view(B, spec.C, 4)
            a.bsums = bsums.data_ptr()
            if has_norm:
                dgb = buf[B * spec.C * 4:].view(2, spec.C)  # rows: d beta | d gamma (contiguous: no copy in AccumulateGrad)
                a.dgb = dgb.data_ptr()
        d_main = torch.empty(main.shape, device=dev, dtype=torch.bfloat16)
        d_add = torch.empty(add.shape, device=dev, dtype=torch.bfloat16) if has_add else None
        dcol = None
        if ctx.dbias_dst is not None:  # bias gradient of the conv that produced `main`, fused into this pass
            dcol = zeros((spec.C,), dev)
        for b0, nb in (_chunks(B, _sample_bytes(main, add, g2, g1)) if two_pass else [(0, 0)]):
            a.b_base, a.b_count = b0, nb
            if two_pass:
                a.d_main, a.d_add, a.dcol = spec.lay_in.view(None), None, None
                _lib.check(L.sb_prep_bwd_reduce(ctypes.byref(a), st), 'sb_prep_bwd_reduce')
            a.d_main = spec.lay_in.view(d_main)
            a.d_add = _ptr(d_add)
            a.dcol = _ptr(dcol)
            _lib.check(L.sb_prep_bwd_apply(ctypes.byref(a), st), 'sb_prep_bwd_apply')
        if dcol is not None:
            ctx.dbias_dst['dbias'] = dcol
        dgamma = dgb[1] if has_norm else None
        dbeta = dgb[0] if has_norm else None
        dsc = bsums[:, :, 2] if has_inj else None
        dsh = bsums[:, :, 3] if has_inj else None
        return None, d_main, d_add, dgamma, dbeta, dsc, dsh, None, None, None


def fused_prep(spec, main, add=None, gamma=None, beta=None, sc=None, sh=None, seed=0, keep=None, dbias_dst=None):
    """`dbias_dst`: the `dbias_src` dict of the ConvSpec that produced `main` when this stage is the ONLY consumer of
    that output: its backward then also emits the conv's bias gradient (column sums of d_main)."""
    out1, out2 = FusedPrep.apply(spec, main, add, gamma, beta, sc, sh, seed, keep, dbias_dst)
    return (out1 if spec.want_out1 else None), out2


# ---------------------------------------------------------------------------------------------------------------
# convolution as tap GEMM
# ---------------------------------------------------------------------------------------------------------------
class ConvSpec:
    """Static description of one convolution lowered to tap GEMMs.

    fwd:   y[B,Ho,Wo,N]      = taps(a [B,Ha,Wa,Ca]; w_fwd [N, T*Ca])         (+bias, relu)
    dgrad: da[B,Ha,Wa,Ca]    = taps^-(dy; w_tr [Ca, T*N])
    wgrad: dw_packed[N,T*Ca] = sum_pix dy (x) a_shifted
    For a Conv2d: w_fwd = pack.fwd, w_tr = pack.tr;  for a ConvTranspose2d the roles are swapped (see sb_pack_conv).
    """

    def __init__(self, name, Ha, Wa, Ca, Ho, Wo, N, taps, relu, transposed=False, need_dgrad=True, bias_tile=1,
                 bwd_mask=False, out_f32=False, flops=0.0):
        self.name = name
        self.flops = flops  # ALGORITHMIC flops per sample of the forward contraction, for rooflines
        self.out_f32 = out_f32  # keep the output in fp32 (tiny deep layers feeding an InstanceNorm over few pixels)
        self.Ha, self.Wa, self.Ca, self.Ho, self.Wo, self.N = Ha, Wa, Ca, Ho, Wo, N
        self.taps, self.relu, self.transposed = taps, relu, transposed
        self.need_dgrad, self.bias_tile = need_dgrad, bias_tile
        # bwd_mask: apply the ReLU mask inside backward (incoming gradients normally arrive already masked by the
        # fused elementwise stage that consumed the output; a residual branch bypasses that stage)
        self.bwd_mask = bwd_mask
        self.neg_taps = [(-dy, -dx) for dy, dx in taps]
        # optional side channel {'dbias': tensor}: a fused consumer of the output (PixelCE) already produced the
        # column sums of the incoming gradient, so backward skips its own pass over dy
        self.dbias_src = None


class PackedWeight:
    """bf16 GEMM operands of one conv-like parameter plus the index maps to (un)pack gradients."""

    def __init__(self, weight, s, Ipad=None, Opad=None, iperm=None, operm=None):
        O, I, KH, KW = weight.shape
        self.O, self.I, self.KH, self.KW, self.s = O, I, KH, KW, s
        self.Ipad, self.Opad = Ipad or I, Opad or O
        self.T, self.PH = (KH // s) * (KW // s), s * s
        dev = weight.device
        self.iperm = torch.tensor(iperm, dtype=torch.int32, device=dev) if iperm is not None else None
        self.operm = torch.tensor(operm, dtype=torch.int32, device=dev) if operm is not None else None
        self.operm_long = self.operm.long() if operm is not None else None            # reference index -> physical index
        self.operm_inv = torch.argsort(self.operm_long) if operm is not None else None  # physical -> reference
        self.fwd = torch.zeros(self.Opad, self.T * self.PH * self.Ipad, device=dev, dtype=torch.bfloat16)
        self.tr = torch.zeros(self.PH * self.Ipad, self.T * self.Opad, device=dev, dtype=torch.bfloat16)
        # packed-gradient mode (set by the Trainer fast path): backward leaves the fp32 gradient in the packed layout in
        # `gpacked` instead of unpacking it into `weight.grad`; the fused Adafactor consumes it and rewrites `fwd`
        self.defer_unpack = False
        self.gpacked = None
        self.grad_slot = None   # optional preallocated (zeroed) fp32 accumulator in the packed layout (DP gradient arena)
        self.on_ready = None    # optional callback fired when this layer's weight gradient has been launched
        self.fresh = False  # operands already match the master weights (written by the optimiser): skip one re-pack
        self.pack(weight)

    def refresh_transposed(self):
        """`fwd` was updated in place (sb_adafactor_conv_packed): re-derive `tr` and mark the operands current."""
        _lib.check(_lib.lib().sb_transpose_taps(self.fwd.data_ptr(), self.tr.data_ptr(), self.Opad, self.PH * self.Ipad,
                                                self.T, _lib.stream_ptr()), 'sb_transpose_taps')
        self.fresh = True

    def pack(self, weight):
        if self.fresh:
            self.fresh = False
            return
        w = weight.detach()
        assert w.dtype == torch.float32 and w.is_contiguous()
        _lib.check(_lib.lib().sb_pack_conv(w.data_ptr(), self.O, self.I, self.KH, self.KW, self.s, self.Ipad,
                                           self.Opad, _ptr(self.iperm), _ptr(self.operm), self.fwd.data_ptr(),
                                           self.tr.data_ptr(), _lib.stream_ptr()), 'sb_pack_conv')

    def unpack_grad(self, g_packed):
        """fp32 gradient in packed 'fwd' layout -> reference layout [O,I,KH,KW]."""
        out = torch.empty(self.O, self.I, self.KH, self.KW, device=g_packed.device, dtype=torch.float32)
        _lib.check(_lib.lib().sb_unpack_conv_grad(g_packed.data_ptr(), self.O, self.I, self.KH, self.KW, self.s,
                                                  self.Ipad, self.Opad, _ptr(self.iperm), _ptr(self.operm),
                                                  out.data_ptr(), _lib.stream_ptr()), 'sb_unpack_conv_grad')
        return out


def colsum(t2d):
    """Column sums of a bf16 [rows, C] matrix (bias gradients) via the per-(b,c) reduction kernel."""
    rows, C = t2d.shape
    spec = PrepSpec(1, rows, C, plain(1, rows, C), plain(1, rows, C))
    return stats_raw(spec, t2d.reshape(1, 1, rows, C))[0, :, 0]


class TapConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, a, weight, bias, pw):
        w_op = pw.tr if spec.transposed else pw.fwd
        bias_f, bias_mod = None, 0
        if bias is not None:
            bias_f = bias.detach().float().contiguous()
            if pw.operm is not None:  # output channels are stored permuted: physical[operm[o]] = reference[o]
                bias_f = bias_f[pw.operm_inv]  # one gather
            if spec.bias_tile > 1:  # N = (phase, channel): the epilogue indexes the bias by column % channels
                bias_mod = spec.N // spec.bias_tile
        B = a.shape[0]
        splits = auto_splits(B, spec.Ho, spec.Wo, spec.N, len(spec.taps) * -(-spec.Ca // 64))
        if splits > 1:
            acc = tap_gemm(a, w_op, spec.taps, spec.Ho, spec.Wo, N=spec.N, splits=splits, tag=spec.name + '.fwd',
                           flops=spec.flops * B)
            y = _bias_act(acc, bias_f, spec.relu, spec.out_f32, bias_mod)
        else:
            y = tap_gemm(a, w_op, spec.taps, spec.Ho, spec.Wo, N=spec.N, bias=bias_f, relu=spec.relu,
                         out_dtype=torch.float32 if spec.out_f32 else torch.bfloat16, tag=spec.name + '.fwd',
                         flops=spec.flops * B, bias_mod=bias_mod)
        ctx.spec, ctx.pw = spec, pw
        ctx.has_bias = bias is not None
        ctx.save_for_backward(a, y if spec.bwd_mask else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, y = ctx.saved_tensors
        if y is not None:
            dy = torch.ops.aten.threshold_backward(dy, y, 0.0)  # dy where y > 0 (ReLU), one kernel
        da, dweight, dbias = _tap_conv_backward(ctx.spec, ctx.pw, a, dy, ctx.needs_input_grad[1], ctx.has_bias)
        return None, da, dweight, dbias, None


def _tap_conv_backward(spec, pw, a, dy, need_da, has_bias):
    """dgrad, wgrad (into the packed layout / the data-parallel arena slot) and bias gradient of one tap convolution."""
    dy = dy.to(dtype=torch.bfloat16, memory_format=torch.contiguous_format)  # (layout + dtype in ONE copy kernel)
    B = a.shape[0]
    da = None
    if spec.need_dgrad and need_da:
        w_op = pw.fwd if spec.transposed else pw.tr
        splits = auto_splits(B, spec.Ha, spec.Wa, spec.Ca, len(spec.taps) * -(-spec.N // 64))
        if splits > 1:
            da = _bias_act(tap_gemm(dy, w_op, spec.neg_taps, spec.Ha, spec.Wa, N=spec.Ca, splits=splits,
                                    tag=spec.name + '.dgrad', flops=spec.flops * B), None, False)
        else:
            da = tap_gemm(dy, w_op, spec.neg_taps, spec.Ha, spec.Wa, N=spec.Ca, tag=spec.name + '.dgrad',
                          flops=spec.flops * B)
    T = len(spec.taps)
    slot = pw.grad_slot if pw.defer_unpack else None
    # bias gradient: (1) handed over by the fused consumer of the output (prep stage / pixel CE), else (2) taken by the
    # weight-gradient kernel from the dy tiles it stages anyway (an extra ones-MMA), else (3) a column-sum pass over dy
    dbias = None
    if has_bias:
        dbias = spec.dbias_src.pop('dbias', None) if spec.dbias_src is not None else None
    mma_bias = None
    if has_bias and dbias is None and not spec.transposed and _lib.lib().sb_tap_wgrad_bias_ok(spec.N, spec.Ca):
        mma_bias = dbias = zeros((spec.N,), a.device)
    if spec.transposed:
        # dW[ci, t, (ph,co)] = sum_pix a[pix, ci] * dy[pix + tap, (ph,co)]
        dwp = slot if slot is not None else zeros((spec.Ca, T * spec.N), a.device)
        tap_wgrad(dy, a, spec.neg_taps, dwp, splits=0,  # 0: the launcher sizes the pixel split for one wave of CTAs
                  tag=spec.name + '.wgrad', flops=spec.flops * B)
    else:
        dwp = slot if slot is not None else zeros((spec.N, T * spec.Ca), a.device)
        tap_wgrad(a, dy, spec.taps, dwp, splits=0, tag=spec.name + '.wgrad', flops=spec.flops * B, dbias=mma_bias)
    if pw.defer_unpack:
        pw.gpacked, dweight = dwp, None
        if pw.on_ready is not None:
            pw.on_ready()
    else:
        dweight = pw.unpack_grad(dwp)
    if has_bias:
        if dbias is None:  # N = (phase, channel) for transposed convs: folding the phases into rows sums them per channel
            dbias = colsum(dy.reshape(-1, spec.N // spec.bias_tile))
        if pw.operm is not None:
            dbias = dbias[pw.operm_long]
    return da, dweight, dbias


def refresh_transposed_all(pws, cache):
    """`fwd` of every PackedWeight in `pws` was rewritten by the optimiser: re-derive all `tr` operands in ONE launch.
    The descriptor table holds device pointers; it lives in `cache` (a dict owned by the caller, e.g. the optimiser's
    per-parameter-set record) and is rebuilt when any operand address changed, so nothing outlives its owner."""
    pws = [pw for pw in pws if pw is not None]
    if not pws:
        return
    key = tuple((pw.fwd.data_ptr(), pw.tr.data_ptr()) for pw in pws)
    rec = cache.get('transpose_table')
    if rec is not None and rec[0] != key:
        rec = None
    if rec is None:
        import numpy as np
        dt = np.dtype([('fwd', np.uint64), ('tr', np.uint64), ('Opad', np.int32), ('PHI', np.int32), ('T', np.int32),
                       ('first_tile', np.int32)])
        assert dt.itemsize == 32
        tab = np.zeros(len(pws), dtype=dt)
        first = 0
        for i, pw in enumerate(pws):
            phi = pw.PH * pw.Ipad
            tab[i] = (pw.fwd.data_ptr(), pw.tr.data_ptr(), pw.Opad, phi, pw.T, first)
            first += -(-phi // 64) * -(-pw.Opad // 64) * pw.T
        dev = torch.from_numpy(tab.view(np.uint8).copy()).to(pws[0].fwd.device)
        rec = cache['transpose_table'] = (key, dev, len(pws), first)
    _, dev, n, total = rec
    _lib.check(_lib.lib().sb_transpose_taps_multi(dev.data_ptr(), n, total, _lib.stream_ptr()), 'sb_transpose_taps_multi')
    for pw in pws:
        pw.fresh = True


SPLIT_FLOOR = True


def auto_splits(B, Ho, Wo, N, KT):
    """Split-K factor for tiny-M layers (deep 3x2 / 6x5 feature maps): enough CTAs to pull the weights at HBM speed
    (one co-resident wave: 148 slots measured a shade faster than two, and halves the fp32 atomics)."""
    from .gemm import choose_box
    bw, bh, bb = choose_box(Wo, Ho, B)
    m_tiles = -(-Wo // bw) * -(-Ho // bh) * -(-B // bb)
    BN = 256 if N % 256 == 0 else 192 if N % 192 == 0 else 128 if N % 128 == 0 else 96 if N <= 96 else 128
    ctas = m_tiles * -(-N // BN)
    if ctas >= 100:
        return 1
    slots = int(os.environ.get('SB_SPLIT_SLOTS', '148'))
    # floor: tiles x splits never exceeds the resident grid, so no CTA of the persistent kernel gets a second tile
    # (ceil gave e.g. 9 tiles x 17 splits = 153 work items on 148 CTAs: five CTAs ran twice as long as the rest).
    # A/B on the replayed step (tools/split_ab.py, ms/step): ceil 7.42 / 7.42, floor 7.35 / 7.56 -- within run-to-run
    # noise, so the rule is chosen on the work-balance argument
    s = slots // ctas if SPLIT_FLOOR else -(-slots // ctas)
    return int(max(1, min(KT, s, 32)))


def _bias_act(acc, bias, relu, keep_f32=False, bias_mod=0):
    """Epilogue of a split-K GEMM (tiny tensors only): fp32 partial sums -> +bias, ReLU, bf16 (or fp32, in place)."""
    N = acc.shape[-1]
    out = acc if keep_f32 else torch.empty(acc.shape, device=acc.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().sb_bias_act(acc.data_ptr(), _ptr(bias), int(bias_mod), int(bool(relu)), out.data_ptr(),
                                      int(keep_f32), acc.numel() // N, N, _lib.stream_ptr()), 'sb_bias_act')
    return out


def tap_conv(spec, a, weight, bias, pw):
    return TapConv.apply(spec, a, weight, bias, pw)


# ---------------------------------------------------------------------------------------------------------------
# recurrent state blend, standardisation, loss
# ---------------------------------------------------------------------------------------------------------------
class GruBlend(torch.autograd.Function):
    """state (fp32, updated IN PLACE) = s*sigmoid(g) + tanh(c)*(1-sigmoid(g)); its bf16 copy goes into xs[..., :64].
    The backward takes the old state from `xs_prev[..., :64]` (the bf16 copy the gate conv just read), so no second
    fp32 state buffer and no commit copy exist."""

    @staticmethod
    def forward(ctx, G, state, xs, xs_prev):
        npix = G.numel() // 128
        assert state.dtype == torch.float32 and state.is_contiguous() and xs_prev.dtype == torch.bfloat16
        _lib.check(_lib.lib().sb_gru_blend(G.data_ptr(), state.data_ptr(), state.data_ptr(), xs.data_ptr(),
                                           xs.shape[-1], npix, _lib.stream_ptr()), 'sb_gru_blend')
        ctx.save_for_backward(G, xs_prev)
        ctx.mark_dirty(xs)
        return xs

    @staticmethod
    def backward(ctx, dxs):
        G, xs_prev = ctx.saved_tensors
        npix = G.numel() // 128
        dxs = dxs.contiguous()
        dG = torch.empty_like(G)
        _lib.check(_lib.lib().sb_gru_blend_bwd(G.data_ptr(), xs_prev.data_ptr(), xs_prev.shape[-1], dxs.data_ptr(),
                                               dxs.shape[-1], dG.data_ptr(), npix, _lib.stream_ptr()),
                   'sb_gru_blend_bwd')
        return dG, None, None, None


def standardize(x, dst=None, coff=0, out_f32=None, dst2=None, coff2=0):
    """x fp32 [B,C,H,W] -> standardised; writes bf16 into dst[..., coff:coff+C] (NHWC), optionally also into
    dst2[..., coff2:coff2+C], and/or fp32 `out_f32` (NCHW)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    B, C, H, W = x.shape
    scratch = torch.empty(B * C * 2, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sb_standardize(x.data_ptr(), B, C, H * W, _ptr(dst), dst.shape[-1] if dst is not None else 0,
                                         coff, _ptr(out_f32), scratch.data_ptr(), _ptr(dst2),
                                         dst2.shape[-1] if dst2 is not None else 0, coff2, _lib.stream_ptr()),
               'sb_standardize')


def frame_feedback(stack, gt_u8, pred_u8, eps, p_noise, seed):
    """In place: stack [B,12,H,W] fp32 <- cat(stack[:, 3:], noisy(gt or pred)/255); see sb_frame_feedback.
    eps: 0-d float device tensor (probability of feeding the ground truth); seed: int64 device tensor or int."""
    B, Cs, H, W = stack.shape
    Cn = gt_u8.shape[1]
    assert stack.dtype == torch.float32 and stack.is_contiguous()
    assert gt_u8.dtype == torch.uint8 and pred_u8.dtype == torch.uint8 and gt_u8.shape == pred_u8.shape
    gt_u8, pred_u8 = gt_u8.contiguous(), pred_u8.contiguous()
    seed_ptr, seed_val = (seed.data_ptr(), 0) if isinstance(seed, torch.Tensor) else (None, int(seed))
    _lib.check(_lib.lib().sb_frame_feedback(stack.data_ptr(), gt_u8.data_ptr(), pred_u8.data_ptr(),
                                            eps.data_ptr() if isinstance(eps, torch.Tensor) else None, float(p_noise),
                                            seed_ptr, seed_val & 0xFFFFFFFFFFFFFFFF, B, Cs, Cn, H * W, _lib.stream_ptr()),
               'sb_frame_feedback')
    return stack


class PixelCE(torch.autograd.Function):
    """Clipped per-pixel CE (simple/trainer.py:134-137) on colour-major bf16 logits [B,H,W,768]; the gradient is
    produced in the same pass and overwrites the logits buffer."""

    @staticmethod
    def forward(ctx, logits, target_u8, clip, want_argmax, unit_grad, dbias_dst=None):
        B, H, W, _ = logits.shape
        HW = H * W
        loss_sum = zeros((1,), logits.device, torch.float64)
        dbias = zeros((768,), logits.device)
        amax = torch.empty(B, 3, H, W, device=logits.device, dtype=torch.uint8) if want_argmax else None
        n = float(B * 3 * HW)
        dl = logits.detach()
        rec = profiling.RECORDER
        ev0 = rec.start() if rec is not None else None
        _lib.check(_lib.lib().sb_pixel_ce(dl.data_ptr(), target_u8.data_ptr(), B, HW, clip, 1.0 / n, dl.data_ptr(),
                                          _ptr(amax), loss_sum.data_ptr(), dbias.data_ptr(), _lib.stream_ptr()),
                   'sb_pixel_ce')
        if rec is not None:  # algorithmic bytes: read logits + write dlogits (bf16) + targets + argmax
            rec.stop('pixel_ce_kernel', 'train', B * HW * (768 * 2 * 2 + 3 + 3), 'byte', ev0)
        ctx.save_for_backward(dl)
        ctx.unit_grad = unit_grad
        if dbias_dst is not None and unit_grad:  # column sums of the gradient, for the logits conv's bias
            dbias_dst['dbias'] = dbias
        loss = (loss_sum / n - clip).float().reshape(())
        if amax is None:
            amax = torch.empty(0, device=logits.device, dtype=torch.uint8)
        ctx.mark_non_differentiable(amax)
        return loss, amax

    @staticmethod
    def backward(ctx, gloss, _):
        (dl,) = ctx.saved_tensors
        # the training loss enters the total with weight 1 (trainer.py:141); honour other weights generically
        if ctx.unit_grad:  # caller guarantees d(total)/d(loss) == 1: skip a full pass over the gradient
            return dl, None, None, None, None, None
        return dl * gloss.to(dl.dtype), None, None, None, None, None


def pixel_argmax(logits):
    """argmax decode of colour-major bf16 logits [B,H,W,768] -> u8 [B,3,H,W] (subproc_vec_env.py:428)."""
    B, H, W, _ = logits.shape
    amax = torch.empty(B, 3, H, W, device=logits.device, dtype=torch.uint8)
    _lib.check(_lib.lib().sb_pixel_ce(logits.data_ptr(), None, B, H * W, 0.0, 0.0, None, amax.data_ptr(), None, None,
                                      _lib.stream_ptr()), 'sb_pixel_ce')
    return amax


def _logits_args(spec, a, pw, bias_p, out):
    from .gemm import _taps, choose_box
    B, H, W, C = a.shape
    bw, bh, bb = choose_box(spec.Wo, spec.Ho, B)
    args = _lib.TapGemmArgs()
    args.a = a.data_ptr(); args.B, args.H, args.W, args.C = B, H, W, C
    args.Ho, args.Wo = spec.Ho, spec.Wo
    args.bw, args.bh, args.bb = bw, bh, bb
    _taps(args, spec.taps)
    args.w = pw.fwd.data_ptr(); args.N = 768; args.ldw = pw.fwd.shape[1]
    args.out = out.data_ptr() if out is not None else None
    args.out_f32, args.ldo = 0, 768
    args.bias = bias_p.data_ptr()
    args.relu, args.splits, args.accumulate, args.bias_mod = 0, 1, 0, 0
    return args


class LogitsCE(torch.autograd.Function):
    """Logits conv (next_frame_predictor.py:356-357) + clipped per-pixel CE + arg-max (trainer.py:130,134-137) in ONE
    tcgen05 kernel (sb_logits_ce): the logits never reach HBM; the forward launch leaves the gradient w.r.t. the logits,
    which the backward feeds to the ordinary dgrad / wgrad kernels.  The loss enters the total with weight 1."""

    @staticmethod
    def forward(ctx, spec, a, weight, bias, pw, target_u8, clip, raw_sum=False):
        B, H, W, _ = a.shape
        n = float(B * 3 * H * W)
        bias_p = bias.detach().float().contiguous()[pw.operm_inv]
        dl = torch.empty(B, H, W, 768, device=a.device, dtype=torch.bfloat16)
        amax = torch.empty(B, 3, H, W, device=a.device, dtype=torch.uint8)
        loss_sum = zeros((1,), a.device, torch.float64)
        args = _logits_args(spec, a, pw, bias_p, dl)
        rec = profiling.RECORDER
        ev0 = rec.start() if rec is not None else None
        _lib.check(_lib.lib().sb_logits_ce(ctypes.byref(args), target_u8.contiguous().data_ptr(), float(clip), 1.0 / n,
                                           amax.data_ptr(), loss_sum.data_ptr(), _lib.stream_ptr()), 'sb_logits_ce')
        if rec is not None:  # algorithmic work of the fused kernel: the GEMM flops, and hf in + dlogits out as bytes
            rec.stop('logits_ce_kernel', 'train', B * H * W * (768 * 2 + a.shape[3] * 2 + 6), 'byte', ev0)
            rec.add_work('tap_gemm_kernel', spec.name + '.fwd+ce', spec.flops * B, 'flop', rec.items[-1])
        ctx.spec, ctx.pw = spec, pw
        ctx.save_for_backward(a, dl)
        ctx.mark_non_differentiable(amax)
        if raw_sum:  # f64 [1]: sum over pixels of max(ce, clip); trainer.TotalLoss turns it into the loss
            return loss_sum, amax
        return (loss_sum / n - clip).float().reshape(()), amax

    @staticmethod
    def backward(ctx, gloss, _):
        a, dl = ctx.saved_tensors
        da, dweight, dbias = _tap_conv_backward(ctx.spec, ctx.pw, a, dl, ctx.needs_input_grad[1], True)
        return None, da, dweight, dbias, None, None, None, None


def logits_argmax(spec, a, bias, pw):
    """Rollout decode (subproc_vec_env.py:428): arg-max frame straight out of the logits GEMM's epilogue."""
    B, H, W, _ = a.shape
    bias_p = bias.detach().float().contiguous()[pw.operm_inv]
    amax = torch.empty(B, 3, H, W, device=a.device, dtype=torch.uint8)
    args = _logits_args(spec, a, pw, bias_p, None)
    _lib.check(_lib.lib().sb_logits_ce(ctypes.byref(args), None, 0.0, 0.0, amax.data_ptr(), None, _lib.stream_ptr()),
               'sb_logits_ce')
    return amax


# ---------------------------------------------------------------------------------------------------------------
# latent-bit predictor
# ---------------------------------------------------------------------------------------------------------------
class MeanAttentionFn(torch.autograd.Function):
    """simple/utils.py:72-88 on the NHWC bf16 conv output [B,H,W,C] (one CTA per sample; see csrc/mean_attention.cu)."""

    @staticmethod
    def forward(ctx, x, we, be, wd, bd):
        B, H, W, C = x.shape
        HW, O = H * W, wd.shape[0]
        x = x.contiguous()
        f = lambda t: t.detach().float().contiguous()
        we2, be_, wd_, bd_ = f(we).reshape(4, C), f(be), f(wd), f(bd)
        y = torch.empty(B, O, device=x.device, dtype=torch.float32)
        s_save = torch.empty(B, 4 * HW, device=x.device, dtype=torch.float32)
        l_save = torch.empty(B, 5 * C, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().sb_mean_attention_fwd(x.data_ptr(), we2.data_ptr(), be_.data_ptr(), wd_.data_ptr(),
                                                    bd_.data_ptr(), B, HW, C, O, y.data_ptr(), s_save.data_ptr(),
                                                    l_save.data_ptr(), _lib.stream_ptr()), 'sb_mean_attention_fwd')
        ctx.save_for_backward(x, we2, wd_, s_save, l_save)
        ctx.we_shape = we.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, we2, wd_, s_save, l_save = ctx.saved_tensors
        B, H, W, C = x.shape
        HW, O = H * W, wd_.shape[0]
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        acc = zeros((4 * C + 4 + O,), x.device)  # dWe [4,C] | dbe [4] | dbd [O]: batch sums accumulated by the kernels
        dwe, dbe, dbd = acc[:4 * C], acc[4 * C:4 * C + 4], acc[4 * C + 4:]
        L = _lib.lib()
        _lib.check(L.sb_mean_attention_bwd(x.data_ptr(), we2.data_ptr(), wd_.data_ptr(), s_save.data_ptr(), dy.data_ptr(),
                                           B, HW, C, O, dx.data_ptr(), dwe.data_ptr(), dbe.data_ptr(), 1,
                                           _lib.stream_ptr()), 'sb_mean_attention_bwd')
        _lib.check(L.sb_colsum_f32(dy.data_ptr(), B, O, O, dbd.data_ptr(), _lib.stream_ptr()), 'sb_colsum_f32')
        return dx, dwe.reshape(ctx.we_shape), dbe, dense(dy, l_save, a_trans=True), dbd


def mean_attention(x_nhwc_bf16, module):
    """module: MeanAttention (input_embedding 1x1 conv to 4 maps + dense on cat(am, m))."""
    return MeanAttentionFn.apply(x_nhwc_bf16, module.input_embedding.weight, module.input_embedding.bias,
                                 module.dense.weight, module.dense.bias)


def lstm_sampler_tables(lstm, dense5, dense4, out=None):
    """Weight-only operands of the sampling kernel (refresh whenever the weights change):
    whhT [128,512], w5T [128,256], E [256,512] = W_ih (W4[:,s] + b4) + b_ih + b_hh, bias = b_ih + b_hh,
    wihT [128,512] = W_ih^T (first-step input gates).  `out`: a dict returned earlier -- rewritten IN PLACE (same device
    addresses: a captured rollout graph keeps reading current values)."""
    with torch.no_grad():
        bias = (lstm.bias_ih + lstm.bias_hh).float()
        emb = dense4.weight.t().float() + dense4.bias.float()  # [256,128]: dense4(one_hot(s)) for every symbol s
        new = dict(whhT=lstm.weight_hh.t().contiguous().float(), w5T=dense5.weight.t().contiguous().float(),
                   E=dense(emb.contiguous(), lstm.weight_ih.detach().float(), b_trans=True, bias=bias), bias=bias,
                   b5=dense5.bias.detach().float().contiguous(), wihT=lstm.weight_ih.t().contiguous().float())
        if out is None or any(out[k].device != v.device for k, v in new.items()):
            return new
        for k, v in new.items():
            out[k].copy_(v)
        return out


def lstm_sample(x0, h0, c0, lstm, dense5, dense4, noise=None, seed=None, tables=None):
    """16-step autoregressive sampling of BitsPredictor (next_frame_predictor.py:109-121) in one kernel launch.
    Returns bits [B,128] in {-1,+1}.  `noise`: explicit Exp(1) draws [B,16,256] (parity mode), else Philox(seed)."""
    B = x0.shape[0]
    f = lambda t: t.detach().float().contiguous()
    tb = tables if tables is not None else lstm_sampler_tables(lstm, dense5, dense4)
    with torch.no_grad():
        G0 = dense(f(x0), tb['wihT'], bias=tb['bias'])  # [B,512]
    h0, c0 = f(h0), f(c0)
    bits = torch.empty(B, 128, device=x0.device, dtype=torch.float32)
    ints = torch.empty(B, 16, device=x0.device, dtype=torch.int32)
    noise = f(noise) if noise is not None else None
    _lib.check(_lib.lib().sb_lstm_sample2(G0.data_ptr(), h0.data_ptr(), c0.data_ptr(), tb['E'].data_ptr(),
                                          tb['whhT'].data_ptr(), tb['w5T'].data_ptr(), tb['b5'].data_ptr(), _ptr(noise),
                                          _ptr(seed), B, h0.stride(0), bits.data_ptr(), ints.data_ptr(),
                                          _lib.stream_ptr()),
               'sb_lstm_sample2')
    return bits, ints


class LstmTeacher(torch.autograd.Function):
    """All 16 teacher-forced LSTMCell steps (next_frame_predictor.py:97-100) in one launch, and their backward through
    time in another.  xw [B,16,512]: input-gate pre-activations (both biases included); returns the outputs [B,16,128]."""

    @staticmethod
    def forward(ctx, xw, h0, c0, w_hh):
        B = xw.shape[0]
        xw, h0, c0 = xw.contiguous(), h0.contiguous(), c0.contiguous()
        whhT = w_hh.detach().t().contiguous()
        hs = torch.empty(B, 16, 128, device=xw.device, dtype=torch.float32)
        act = torch.empty(B, 16, 512, device=xw.device, dtype=torch.float32)
        cs = torch.empty(B, 16, 128, device=xw.device, dtype=torch.float32)
        _lib.check(_lib.lib().sb_lstm_teacher_fwd(xw.data_ptr(), h0.data_ptr(), c0.data_ptr(), whhT.data_ptr(), B, 128,
                                                  hs.data_ptr(), act.data_ptr(), cs.data_ptr(), _lib.stream_ptr()),
                   'sb_lstm_teacher_fwd')
        ctx.save_for_backward(act, cs, hs, h0, c0, w_hh)
        return hs

    @staticmethod
    def backward(ctx, dhs):
        act, cs, hs, h0, c0, w_hh = ctx.saved_tensors
        B = act.shape[0]
        dhs = dhs.contiguous()
        whh = w_hh.detach().contiguous()
        dpre = torch.empty_like(act)
        dh0 = torch.empty_like(h0)
        dc0 = torch.empty_like(c0)
        _lib.check(_lib.lib().sb_lstm_teacher_bwd(dhs.data_ptr(), act.data_ptr(), cs.data_ptr(), c0.data_ptr(),
                                                  whh.data_ptr(), B, 128, dpre.data_ptr(), dh0.data_ptr(), dc0.data_ptr(),
                                                  _lib.stream_ptr()), 'sb_lstm_teacher_bwd')
        h_prev = torch.cat((h0.unsqueeze(1), hs[:, :-1]), dim=1).reshape(-1, 128)
        dw_hh = dense(dpre.reshape(-1, 512), h_prev, a_trans=True)  # [512,128]
        return dpre, dh0, dc0, dw_hh
