"""tcgen05 tap-GEMM / wgrad kernels vs a plain PyTorch fp32 reference of the same contraction (bf16 inputs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device('cuda:0')


@pytest.mark.parametrize('M,K,N', [(300, 256, 128), (128, 64, 96), (1000, 384, 192), (517, 512, 256), (64, 1024, 768)])
def test_plain_gemm(M, K, N):
    from simple_b200.gemm import tap_gemm
    torch.manual_seed(0)
    a = torch.randn(1, 1, M, K, device=_dev()).bfloat16()
    w = torch.randn(N, K, device=_dev()).bfloat16()
    bias = torch.randn(N, device=_dev())
    out = tap_gemm(a, w, [(0, 0)], 1, M, bias=bias, relu=True, out_dtype=torch.float32)
    ref = F.relu(a.float().reshape(M, K) @ w.float().t() + bias)
    torch.cuda.synchronize()
    assert torch.allclose(out.reshape(M, N), ref, rtol=1e-3, atol=1e-2), float((out.reshape(M, N) - ref).abs().max())
    out16 = tap_gemm(a, w, [(0, 0)], 1, M, bias=bias, relu=False)
    ref16 = (a.float().reshape(M, K) @ w.float().t() + bias)
    assert torch.allclose(out16.float().reshape(M, N), ref16, rtol=1e-2, atol=1e-1)


@pytest.mark.parametrize('splits', [1, 3])
def test_conv3x3(splits):
    from simple_b200.gemm import tap_gemm
    torch.manual_seed(1)
    B, H, W, C, N = 3, 13, 10, 128, 256
    x = torch.randn(B, H, W, C, device=_dev()).bfloat16()
    wt = (torch.randn(N, C, 3, 3, device=_dev()) * 0.05).bfloat16()
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    wp = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    out = tap_gemm(x, wp, taps, H, W, splits=splits, out_dtype=torch.float32)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-2), float((out - ref).abs().max())


def test_conv4x4s2_via_s2d():
    """4x4 stride-2 pad-1 conv == 2x2 stride-1 taps over the padded space-to-depth input."""
    from simple_b200.gemm import tap_gemm
    torch.manual_seed(2)
    B, H, W, C, N = 2, 105, 80, 96, 192
    Ho, Wo = 52, 40
    x = torch.randn(B, C, H, W, device=_dev()).bfloat16()
    wt = (torch.randn(N, C, 4, 4, device=_dev()) * 0.05).bfloat16()
    xp = F.pad(x, (1, 1, 1, 1))  # [B,C,107,82]
    H2, W2 = Ho + 1, Wo + 1
    xp = xp[:, :, :2 * H2, :2 * W2]
    s2d = xp.reshape(B, C, H2, 2, W2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H2, W2, 4 * C).contiguous()
    # packed weights [N, (dy,dx),(py,px),c] with ky = 2dy+py
    wp = wt.reshape(N, C, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(N, 16 * C).contiguous()  # o,dy,dx,py,px,c
    taps = [(dy, dx) for dy in range(2) for dx in range(2)]
    out = tap_gemm(s2d, wp, taps, Ho, Wo, out_dtype=torch.float32)
    ref = F.conv2d(x.float(), wt.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, rtol=1e-3, atol=2e-2), float((out - ref).abs().max())


@pytest.mark.parametrize('C,N,splits', [(128, 256, 1), (64, 96, 2), (384, 192, 3)])
def test_wgrad(C, N, splits):
    from simple_b200.gemm import tap_wgrad
    torch.manual_seed(3)
    B, H, W = 5, 13, 10
    x = torch.randn(B, H, W, C, device=_dev()).bfloat16()
    dy = torch.randn(B, H, W, N, device=_dev()).bfloat16()
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    dw = torch.zeros(N, 9 * C, device=_dev())
    # bias gradient by the extra ones-MMA of the weight-gradient kernel (normal tile orientation only)
    from simple_b200 import _lib
    dbias = torch.zeros(N, device=_dev()) if _lib.lib().sb_tap_wgrad_bias_ok(N, C) else None
    tap_wgrad(x, dy, taps, dw, splits=splits, dbias=dbias)
    if dbias is not None:
        torch.cuda.synchronize()
        ref_b = dy.float().sum(dim=(0, 1, 2))
        assert float((dbias - ref_b).abs().max()) <= 1e-3 * float(ref_b.abs().max()) + 1e-3, (dbias[:4], ref_b[:4])
    xn = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.zeros(N, C, 3, 3, device=_dev(), requires_grad=True)
    y = F.conv2d(xn, wref, padding=1)
    (g,) = torch.autograd.grad(y, wref, dy.float().permute(0, 3, 1, 2))
    ref = g.permute(0, 2, 3, 1).reshape(N, 9 * C)
    torch.cuda.synchronize()
    err = float((dw - ref).abs().max())
    assert torch.allclose(dw, ref, rtol=1e-3, atol=5e-2), err


@pytest.mark.parametrize('B,Ho,Wo,C,N,kind', [(9, 52, 40, 96, 192, 's2d'), (5, 105, 80, 96, 768, '1x1'), (5, 105, 80, 768, 96, '1x1'),
                                               (7, 105, 80, 128, 128, '3x3'), (33, 26, 20, 192, 384, 's2d')])
def test_cta_pair_kernel_matches_single_cta_and_reference(B, Ho, Wo, C, N, kind):
    """Large bf16-output layers run on the CTA-pair (tcgen05 cta_group::2, M = 256) kernel.  It must agree with the
    single-CTA kernel bit for bit (same products, same fp32 accumulation order per output) and with torch; odd M-tile
    counts exercise the clipped last half-tile, bias + ReLU the epilogue."""
    from simple_b200 import _lib
    from simple_b200.gemm import tap_gemm
    torch.manual_seed(5)
    if kind == 's2d':
        a = torch.randn(B, Ho + 1, Wo + 1, 4 * C, device=_dev()).bfloat16()
        taps = [(dy, dx) for dy in range(2) for dx in range(2)]
        Ca = 4 * C
    elif kind == '3x3':
        a = torch.randn(B, Ho, Wo, C, device=_dev()).bfloat16()
        taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
        Ca = C
    else:
        a = torch.randn(B, Ho, Wo, C, device=_dev()).bfloat16()
        taps = [(0, 0)]
        Ca = C
    w = (torch.randn(N, len(taps) * Ca, device=_dev()) * 0.05).bfloat16()
    bias = torch.randn(N, device=_dev())
    L = _lib.lib()
    prev = L.sb_set_pair_mode(1)
    try:
        n0 = L.sb_launch_count()
        out_pair = tap_gemm(a, w, taps, Ho, Wo, N=N, bias=bias, relu=True)
        L.sb_set_pair_mode(0)
        out_single = tap_gemm(a, w, taps, Ho, Wo, N=N, bias=bias, relu=True)
    finally:
        L.sb_set_pair_mode(prev)
    torch.cuda.synchronize()
    assert torch.equal(out_pair, out_single), float((out_pair.float() - out_single.float()).abs().max())
    # torch reference of the same tap contraction, one sample at a time (fp32)
    for b in (0, B - 1):
        acc = torch.zeros(Ho, Wo, N, device=_dev())
        for t, (dy, dx) in enumerate(taps):
            src = torch.zeros(Ho, Wo, Ca, device=_dev())
            ys, ye = max(0, -dy), min(Ho, a.shape[1] - dy)
            xs, xe = max(0, -dx), min(Wo, a.shape[2] - dx)
            src[ys:ye, xs:xe] = a[b, ys + dy:ye + dy, xs + dx:xe + dx].float()
            acc += src @ w[:, t * Ca:(t + 1) * Ca].float().t()
        ref = F.relu(acc + bias)
        got = out_pair[b].float()
        assert torch.allclose(got, ref, rtol=2e-2, atol=5e-2), float((got - ref).abs().max())


def test_ragged_channel_operand_c80():
    """The gate / embedding operand has 80 channels (76 used): its second 64-channel TMA box is zero-filled past channel
    80 and only the valid K=16 steps run; wgrad runs the MMA at N = 80; dgrad writes an 80-channel output through the
    clipped TMA store.  All three against torch."""
    from simple_b200.gemm import tap_gemm, tap_wgrad
    torch.manual_seed(9)
    B, H, W, C, N = 6, 21, 16, 80, 128
    x = torch.randn(B, H, W, C, device=_dev()).bfloat16()
    wt = (torch.randn(N, C, 3, 3, device=_dev()) * 0.05).bfloat16()
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    wp = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    out = tap_gemm(x, wp, taps, H, W, out_dtype=torch.bfloat16)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.allclose(out.float(), ref, rtol=2e-2, atol=5e-2), float((out.float() - ref).abs().max())
    # wgrad, A-channel side = 80
    dy = torch.randn(B, H, W, N, device=_dev()).bfloat16()
    dw = torch.zeros(N, 9 * C, device=_dev())
    tap_wgrad(x, dy, taps, dw, splits=2)
    wref = torch.zeros(N, C, 3, 3, device=_dev(), requires_grad=True)
    (g,) = torch.autograd.grad(F.conv2d(x.float().permute(0, 3, 1, 2), wref, padding=1), wref, dy.float().permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    assert torch.allclose(dw, g.permute(0, 2, 3, 1).reshape(N, 9 * C), rtol=1e-3, atol=5e-2)
    # 1x1 "dgrad" with an 80-channel output (ragged last 32-column chunk, no bias)
    w2 = (torch.randn(C, 96, device=_dev()) * 0.1).bfloat16()     # [N_out = 80, K = 96]
    y = torch.randn(B, H, W, 96, device=_dev()).bfloat16()
    dx = tap_gemm(y, w2, [(0, 0)], H, W, N=C)
    torch.cuda.synchronize()
    ref2 = y.float() @ w2.float().t()
    assert dx.shape == (B, H, W, C) and torch.allclose(dx.float(), ref2, rtol=2e-2, atol=5e-2)


def test_wgrad_swapped_orientation_and_bn192():
    """A 192 x 384 gradient (down0 / up4 of the world model) is tiled 384 x 192 -- A channels on the 128-row side, dY
    channels on a 192-column tile -- instead of a 44 %-padded 256 x 512; plus automatic pixel splits.  Against torch."""
    from simple_b200.gemm import tap_wgrad
    torch.manual_seed(21)
    B, H, W, C, N = 4, 14, 10, 384, 192   # A: [B,H+1,W+1,C] (2x2 taps over a padded input), dY: [B,H,W,N]
    a = torch.randn(B, H + 1, W + 1, C, device=_dev()).bfloat16()
    dy = torch.randn(B, H, W, N, device=_dev()).bfloat16()
    taps = [(dy_, dx_) for dy_ in range(2) for dx_ in range(2)]
    dw = torch.zeros(N, 4 * C, device=_dev())
    tap_wgrad(a, dy, taps, dw, splits=0)
    torch.cuda.synchronize()
    ref = torch.zeros(N, 4 * C, device=_dev())
    for t, (ty, tx) in enumerate(taps):
        at = a[:, ty:ty + H, tx:tx + W].float().reshape(-1, C)
        ref[:, t * C:(t + 1) * C] = dy.float().reshape(-1, N).t() @ at
    assert torch.allclose(dw, ref, rtol=1e-3, atol=5e-2), float((dw - ref).abs().max())


@pytest.mark.parametrize('C,N,with_bias', [(256, 256, False), (768, 768, False), (96, 768, True), (128, 384, True), (384, 192, False)])
def test_wgrad_two_m_tiles_per_cta_matches_one(C, N, with_bias):
    """MT = 2 (two dY tiles per pipeline stage and CTA) against the one-M-tile kernel and torch, with and without the
    ones-MMA bias gradient; (96, 768) is the logits layer's shape, (384, 192) runs in the swapped orientation."""
    from simple_b200 import _lib
    from simple_b200.gemm import tap_wgrad
    torch.manual_seed(5)
    B, H, W = 6, 13, 10
    x = torch.randn(B, H, W, C, device=_dev()).bfloat16()
    dy = torch.randn(B, H, W, N, device=_dev()).bfloat16()
    taps = [(0, 0), (0, 1), (1, 0), (1, 1)]
    L = _lib.lib()
    res = []
    prev = L.sb_set_wgrad_mt2(1)
    try:
        for mode in (1, 0):
            L.sb_set_wgrad_mt2(mode)
            dw = torch.zeros(N, 4 * C, device=_dev())
            db = torch.zeros(N, device=_dev()) if with_bias and L.sb_tap_wgrad_bias_ok(N, C) else None
            tap_wgrad(x, dy, taps, dw, splits=0, dbias=db)
            res.append((dw, db))
    finally:
        L.sb_set_wgrad_mt2(prev)
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    xp = F.pad(xn, (0, 1, 0, 1))
    wref = torch.zeros(N, C, 2, 2, device=_dev(), requires_grad=True)
    (g,) = torch.autograd.grad(F.conv2d(xp, wref), wref, dy.float().permute(0, 3, 1, 2))
    ref = g.permute(0, 2, 3, 1).reshape(N, 4 * C)
    scale = float(ref.abs().max())
    assert float((res[0][0] - res[1][0]).abs().max()) <= 2e-5 * scale
    assert float((res[0][0] - ref).abs().max()) <= 2e-3 * scale
    if res[0][1] is not None:
        ref_b = dy.float().sum(dim=(0, 1, 2))
        for _, db in res:
            assert float((db - ref_b).abs().max()) <= 1e-3 * float(ref_b.abs().max()) + 1e-3


def test_wgrad_tap_groups_match_single_tap_kernel_and_torch():
    """The gate conv's weight gradient (3x3 taps, 128 x 80 per tap, half a million pixels) runs with THREE taps per CTA:
    one dY tile per pipeline stage feeds three shifted A tiles and three TMEM accumulators.  Same result as the
    one-tap-per-CTA kernel (up to the order of the fp32 split reductions) and as torch; the image border exercises the
    TMA zero fill of every shifted tile."""
    from simple_b200 import _lib
    from simple_b200.gemm import tap_wgrad
    torch.manual_seed(31)
    B, H, W, C, N = 8, 105, 80, 80, 128
    x = torch.randn(B, H, W, C, device=_dev()).bfloat16()
    x[..., 76:] = 0
    dy = (torch.randn(B, H, W, N, device=_dev()) * 0.1).bfloat16()
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    L = _lib.lib()
    dw_g = torch.zeros(N, 9 * C, device=_dev())
    dw_1 = torch.zeros(N, 9 * C, device=_dev())
    db_g = torch.zeros(N, device=_dev())
    db_1 = torch.zeros(N, device=_dev())
    prev = L.sb_set_wgrad_tap_groups(1)
    try:
        tap_wgrad(x, dy, taps, dw_g, splits=0, dbias=db_g)
        L.sb_set_wgrad_tap_groups(0)
        tap_wgrad(x, dy, taps, dw_1, splits=0, dbias=db_1)
    finally:
        L.sb_set_wgrad_tap_groups(prev)
    wref = torch.zeros(N, C, 3, 3, device=_dev(), requires_grad=True)
    (g,) = torch.autograd.grad(F.conv2d(x.float().permute(0, 3, 1, 2), wref, padding=1), wref,
                               dy.float().permute(0, 3, 1, 2))
    ref = g.permute(0, 2, 3, 1).reshape(N, 9 * C)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((dw_g - dw_1).abs().max()) <= 2e-5 * scale
    assert float((dw_g - ref).abs().max()) <= 2e-3 * scale, float((dw_g - ref).abs().max()) / scale
    ref_b = dy.float().sum(dim=(0, 1, 2))
    for db in (db_g, db_1):  # bias gradient from the ones-MMA, in both kernel variants
        assert float((db - ref_b).abs().max()) <= 1e-3 * float(ref_b.abs().max()) + 1e-3
