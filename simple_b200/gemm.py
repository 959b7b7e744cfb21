"""Host-side wrappers of the tcgen05 tap-GEMM kernels (sb_tap_gemm / sb_tap_wgrad)."""
import ctypes
import functools

import torch

from . import _lib, profiling


@functools.lru_cache(maxsize=None)
def choose_box(W, H, B, rows=128, exact=False):
    """Pixel box (bw, bh, bb) with bw*bh*bb <= rows (== rows if `exact`) that wastes the fewest tensor-core rows."""
    best = None
    for bw in range(1, min(W, rows) + 1) if not exact else [1, 2, 4, 8, 16, 32, 64, 128]:
        if bw > rows or (exact and bw > 2 * W):
            continue
        for bh in (range(1, min(H, rows // bw) + 1) if not exact else [1, 2, 4, 8, 16, 32, 64, 128]):
            if bw * bh > rows or (exact and bh > 2 * H):
                continue
            bb = rows // (bw * bh)
            if exact:
                if bw * bh * bb != rows:
                    continue
            else:
                bb = max(1, min(bb, B))
            if bb > 256:
                continue
            tiles = -(-W // bw) * -(-H // bh) * -(-B // bb)
            key = (tiles, bb, abs(bw - bh))
            if best is None or key < best[0]:
                best = (key, (bw, bh, bb))
    return best[1]


def _taps(args, taps):
    args.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        args.tap_dy[i] = dy
        args.tap_dx[i] = dx


def tap_gemm(a, w, taps, Ho, Wo, N=None, out=None, out_dtype=torch.bfloat16, bias=None, relu=False, splits=1,
             accumulate=False, box=None, tag=None, flops=0.0, bias_mod=0):
    """out[b,y,x,n] (+)= sum_t sum_c a[b,y+dy_t,x+dx_t,c] * w[n, t*C+c].  a: bf16 [B,H,W,C]; w: bf16 [N, >=T*C]."""
    assert a.is_cuda and a.dtype == torch.bfloat16 and a.is_contiguous() and a.dim() == 4
    assert w.dtype == torch.bfloat16 and w.is_contiguous() and w.dim() == 2
    B, H, W, C = a.shape
    N = N or w.shape[0]
    atomic = splits > 1 or accumulate
    if out is None:
        dt = torch.float32 if atomic else out_dtype
        if atomic:
            from .ops import zeros
            out = zeros((B, Ho, Wo, N), a.device, dt)
        else:
            out = torch.empty((B, Ho, Wo, N), device=a.device, dtype=dt)
    assert out.is_contiguous() and out.shape[:3] == (B, Ho, Wo)
    bw, bh, bb = box or choose_box(Wo, Ho, B)
    args = _lib.TapGemmArgs()
    args.a = a.data_ptr(); args.B, args.H, args.W, args.C = B, H, W, C
    args.Ho, args.Wo = Ho, Wo
    args.bw, args.bh, args.bb = bw, bh, bb
    _taps(args, taps)
    args.w = w.data_ptr(); args.N = N; args.ldw = w.shape[1]
    args.out = out.data_ptr(); args.out_f32 = int(out.dtype == torch.float32); args.ldo = out.shape[3]
    args.bias = bias.data_ptr() if bias is not None else None
    args.bias_mod = int(bias_mod)
    args.relu = int(relu); args.splits = splits; args.accumulate = int(accumulate)
    rec = profiling.RECORDER
    ev0 = rec.start() if rec is not None else None
    _lib.check(_lib.lib().sb_tap_gemm(ctypes.byref(args), _lib.stream_ptr()), 'sb_tap_gemm')
    if rec is not None:
        rec.stop('tap_gemm_kernel', tag, flops, 'flop', ev0)
    return out


def tap_wgrad(a, dy, taps, dw, splits=1, box=None, tag=None, flops=0.0, dbias=None):
    """dw[n, t*C+c] += sum_{b,y,x} dy[b,y,x,n] * a[b,y+dy_t,x+dx_t,c].  dw: fp32 [N, >=T*C] (accumulated)."""
    assert a.dtype == torch.bfloat16 and a.is_contiguous() and dy.dtype == torch.bfloat16 and dy.is_contiguous()
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    B, H, W, C = a.shape
    B2, Ho, Wo, N = dy.shape
    assert B == B2 and dw.shape[0] == N
    bw, bh, bb = box or choose_box(Wo, Ho, B, rows=64, exact=True)
    args = _lib.TapWgradArgs()
    args.a = a.data_ptr(); args.B, args.H, args.W, args.C = B, H, W, C
    args.dy = dy.data_ptr(); args.Ho, args.Wo, args.N = Ho, Wo, N
    args.bw, args.bh, args.bb = bw, bh, bb
    _taps(args, taps)
    args.dw = dw.data_ptr(); args.ldw = dw.shape[1]; args.splits = splits
    args.dbias = dbias.data_ptr() if dbias is not None else None
    rec = profiling.RECORDER
    ev0 = rec.start() if rec is not None else None
    _lib.check(_lib.lib().sb_tap_wgrad(ctypes.byref(args), _lib.stream_ptr()), 'sb_tap_wgrad')
    if rec is not None:
        rec.stop('tap_wgrad_kernel', tag, flops, 'flop', ev0)
    return dw
