"""Full-size (BASELINE configs[1]: batch 64, 105x80 frames) checks of the hot kernels through size-independent
properties: exact linearity under power-of-two scaling, pair-vs-single kernel identity, checksum-of-checksums for the
loss gradient, argmax / loss against a chunked torch fp32 evaluation.  The oracle would take minutes at this size; these
properties hold for any input, so random data at the real shapes is enough."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

B, H, W = 64, 105, 80


def _dev():
    return torch.device('cuda:0')


@pytest.mark.parametrize('layer', ['gate3x3_c80', 'logits1x1', 'down0_s2d'])
def test_tap_gemm_full_size_linearity_and_kernel_identity(layer):
    """conv(4x) == 4 conv(x) and conv(x; 0.5 w) == 0.5 conv(x; w) BIT-EXACTLY (scaling by powers of two commutes with
    bf16 rounding and fp32 accumulation), at the real layer shapes; the CTA-pair kernel equals the single-CTA kernel."""
    from simple_b200 import _lib
    from simple_b200.gemm import tap_gemm
    torch.manual_seed(11)
    if layer == 'gate3x3_c80':
        a = torch.randn(B, H, W, 80, device=_dev()).bfloat16()
        a[..., 76:] = 0
        taps, Ho, Wo, N = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)], H, W, 128
    elif layer == 'logits1x1':
        a = torch.randn(B, H, W, 96, device=_dev()).bfloat16()
        taps, Ho, Wo, N = [(0, 0)], H, W, 768
    else:
        a = torch.randn(B, 53, 41, 4 * 96, device=_dev()).bfloat16()
        taps, Ho, Wo, N = [(dy, dx) for dy in range(2) for dx in range(2)], 52, 40, 192
    K = len(taps) * a.shape[-1]
    w = (torch.randn(N, K, device=_dev()) * 0.05).bfloat16()
    L = _lib.lib()
    y = tap_gemm(a, w, taps, Ho, Wo, N=N)
    y4 = tap_gemm(a * 4, w, taps, Ho, Wo, N=N)
    yh = tap_gemm(a, w * 0.5, taps, Ho, Wo, N=N)
    prev = L.sb_set_pair_mode(0)
    try:
        ys = tap_gemm(a, w, taps, Ho, Wo, N=N)
    finally:
        L.sb_set_pair_mode(prev)
    torch.cuda.synchronize()
    assert y.shape == (B, Ho, Wo, N) and bool(torch.isfinite(y.float()).all())
    assert torch.equal(y4, y * 4)
    assert torch.equal(yh, y * 0.5)
    assert torch.equal(ys, y)
    # spot check against torch fp32 on the last sample (catches batch-offset errors past 2^31 bytes)
    acc = torch.zeros(Ho, Wo, N, device=_dev())
    Ca = a.shape[-1]
    for t, (dy, dx) in enumerate(taps):
        src = torch.zeros(Ho, Wo, Ca, device=_dev())
        y0, y1 = max(0, -dy), min(Ho, a.shape[1] - dy)
        x0, x1 = max(0, -dx), min(Wo, a.shape[2] - dx)
        src[y0:y1, x0:x1] = a[B - 1, y0 + dy:y1 + dy, x0 + dx:x1 + dx].float()
        acc += src @ w[:, t * Ca:(t + 1) * Ca].float().t()
    assert torch.allclose(y[B - 1].float(), acc, rtol=2e-2, atol=5e-2), float((y[B - 1].float() - acc).abs().max())


def test_wgrad_full_size_linearity():
    """Weight gradient of the logits layer at full size (537 600 pixels contracted, HBM-bound): scaling dY by 2 scales
    every product by exactly 2; the fp32 atomics of the pixel splits add the partial sums in a different order each
    run, so linearity is checked to 1e-5 relative.  64 gradient rows against an fp64 contraction pin the value."""
    from simple_b200.gemm import tap_wgrad
    torch.manual_seed(12)
    x = torch.randn(B, H, W, 96, device=_dev()).bfloat16()
    dy = (torch.randn(B, H, W, 768, device=_dev()) * 0.01).bfloat16()
    dw = torch.zeros(768, 96, device=_dev())
    dw2 = torch.zeros(768, 96, device=_dev())
    tap_wgrad(x, dy, [(0, 0)], dw, splits=0)
    tap_wgrad(x, dy * 2, [(0, 0)], dw2, splits=0)
    torch.cuda.synchronize()
    scale = float(dw.abs().max())
    assert float((dw2 - 2 * dw).abs().max()) <= 1e-5 * 2 * scale + 1e-6
    rows = torch.arange(0, 768, 12, device=_dev())
    ref = dy.reshape(-1, 768)[:, rows].double().t() @ x.reshape(-1, 96).double()
    got = dw[rows].double()
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max()) + 1e-4


def test_pixel_ce_full_size_checksums_argmax_and_loss():
    """B=64 x 8400 pixels x 3 colours x 256 logits (826 MB).  (1) the gradient of every group sums to ~0 and the fused
    bias gradient equals the column sums of the gradient (checksum of checksums); (2) argmax is bit-exact vs torch on
    the same bf16 logits; (3) the clipped loss equals a chunked fp32 torch evaluation."""
    from simple_b200 import ops
    torch.manual_seed(13)
    HW = H * W
    logits = (torch.randn(B, H, W, 768, device=_dev()) * 2).bfloat16()  # colour-major: channel = colour * 256 + value
    lview = logits.view(B, HW, 3, 256)
    target = torch.randint(0, 256, (B, 3, H, W), device=_dev(), dtype=torch.uint8)
    ref_loss = torch.zeros((), device=_dev(), dtype=torch.float64)
    ref_arg = torch.empty(B, 3, HW, device=_dev(), dtype=torch.uint8)
    ref_db = torch.zeros(3, 256, device=_dev(), dtype=torch.float64)
    for b in range(B):  # chunked: fp32 copies of one sample at a time
        lf = lview[b].float()  # [HW,3,256]
        lp = F.log_softmax(lf, dim=-1)
        t = target[b].reshape(3, HW).t().long()  # [HW,3]
        ce = -lp.gather(-1, t[..., None])[..., 0]
        ref_loss += ce.clamp(min=0.03).double().sum()
        ref_arg[b] = lf.argmax(-1).t().to(torch.uint8)
        gb = lp.exp()
        gb.scatter_add_(-1, t[..., None], -torch.ones_like(gb[..., :1]))
        ref_db += (gb * (ce >= 0.03)[..., None]).double().sum(0)  # clipped groups have no gradient
    ref_loss = float(ref_loss / (B * 3 * HW) - 0.03)
    lg = logits.clone().requires_grad_(True)
    side = {}
    loss, amax = ops.PixelCE.apply(lg, target, 0.03, True, True, side)  # gradient (x 1/n) overwrites lg in place
    torch.cuda.synchronize()
    n = B * 3 * HW
    assert abs(float(loss.detach()) - ref_loss) <= 1e-3 * abs(ref_loss) + 1e-4, (float(loss.detach()), ref_loss)
    assert torch.equal(amax.reshape(B, 3, HW), ref_arg)
    col = torch.zeros(768, device=_dev(), dtype=torch.float64)
    gross = torch.zeros(768, device=_dev(), dtype=torch.float64)
    worst = 0.0
    for b in range(B):
        g = lg.detach()[b].float().reshape(-1, 256)
        worst = max(worst, float(g.sum(-1).abs().max()))
        col += g.reshape(-1, 768).double().sum(0)
        gross += g.reshape(-1, 768).double().abs().sum(0)
    # softmax - onehot sums to zero per group; what is left is the bf16 rounding of 256 terms of size <= 1/n
    assert worst <= 0.01 / n, worst * n
    # the fused bias gradient is accumulated in fp32 BEFORE the bf16 rounding of the stored gradient: it must match the
    # fp32 torch evaluation closely, and the column sums of the stored bf16 gradient to within that rounding (the
    # ~2100 one-hot entries of a column all sit just below -1/n and round the same way: 2^-9 of the gross mass)
    dbias = side['dbias'].double()
    ref_db = (ref_db / n).reshape(768)
    assert float((dbias - ref_db).abs().max()) <= 2e-3 * float(ref_db.abs().max()), (dbias - ref_db).abs().max()
    assert bool(((dbias - col).abs() <= gross * 2.0 ** -9).all())
