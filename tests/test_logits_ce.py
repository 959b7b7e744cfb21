"""The logits GEMM fused with the softmax cross-entropy (sb_logits_ce: CE in the epilogue of the tcgen05 CTA-pair kernel)
against the two-kernel path it replaces (sb_tap_gemm -> bf16 logits in HBM -> sb_pixel_ce): same loss, IDENTICAL arg-max
frames (ties included), gradient tile equal up to the summation order of the fp32 exp-sums, and the same parameter /
input gradients through the autograd wrappers."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _setup(B, seed=0, peaked=False):
    from oracle.wm_oracle import default_config
    from simple_b200.next_frame_predictor import NextFramePredictor
    cfg = default_config(device=DEV, batch_size=B)
    torch.manual_seed(seed)
    model = NextFramePredictor(cfg, 3).to(DEV)
    if peaked:
        with torch.no_grad():
            model.logits.weight.mul_(20.0)
            model.logits.bias.normal_(0, 2.0)
    model._plan = model._build_plan(torch.device(DEV))
    model._packed = model._build_packed()
    g = torch.Generator().manual_seed(seed + 1)
    hf = torch.randn(B, 105, 80, 96, generator=g).to(DEV).bfloat16()
    target = torch.randint(0, 256, (B, 3, 105, 80), generator=g, dtype=torch.uint8).to(DEV)
    return cfg, model, hf, target


@pytest.mark.parametrize('B,peaked', [(2, False), (3, True), (16, False)])
def test_fused_logits_ce_matches_gemm_then_pixel_ce(B, peaked):
    from simple_b200 import ops
    cfg, model, hf, target = _setup(B, peaked=peaked)
    spec, pw, m = model._plan['logits'], model._packed['logits'], model.logits
    clip = 0.03
    # ---- two-kernel path
    hf_a = hf.clone().requires_grad_(True)
    model.zero_grad(set_to_none=True)
    logits = ops.tap_conv(spec, hf_a, m.weight, m.bias, pw)
    logits_copy = logits.detach().clone()
    spec.dbias_src.clear()
    loss_a, amax_a = ops.PixelCE.apply(logits, target, clip, True, True, spec.dbias_src)
    loss_a.backward()
    dl_a = logits.detach().clone()  # PixelCE overwrote the logits with their gradient
    ga = (hf_a.grad.clone(), m.weight.grad.clone(), m.bias.grad.clone())
    # ---- fused kernel
    hf_b = hf.clone().requires_grad_(True)
    model.zero_grad(set_to_none=True)
    loss_b, amax_b = ops.LogitsCE.apply(spec, hf_b, m.weight, m.bias, pw, target, clip)
    loss_b.backward()
    gb = (hf_b.grad.clone(), m.weight.grad.clone(), m.bias.grad.clone())
    torch.cuda.synchronize()
    assert torch.equal(amax_a, amax_b), float((amax_a != amax_b).float().mean())
    assert torch.equal(amax_a.long().cpu(),
                       logits_copy.reshape(B, 105, 80, 3, 256).permute(0, 3, 1, 2, 4).float().argmax(-1).cpu())
    assert abs(float(loss_a) - float(loss_b)) < 2e-6 * max(1.0, abs(float(loss_a))), (float(loss_a), float(loss_b))
    # gradient tile of the fused launch (saved for backward): re-run the raw kernel to look at it
    dl_b = torch.empty(B, 105, 80, 768, device=DEV, dtype=torch.bfloat16)
    import ctypes
    from simple_b200 import _lib
    bias_p = m.bias.detach().float().contiguous()[pw.operm_inv]
    args = ops._logits_args(spec, hf, pw, bias_p, dl_b)
    loss_sum = torch.zeros(1, device=DEV, dtype=torch.float64)
    am = torch.empty(B, 3, 105, 80, device=DEV, dtype=torch.uint8)
    n = float(B * 3 * 105 * 80)
    _lib.check(_lib.lib().sb_logits_ce(ctypes.byref(args), target.data_ptr(), clip, 1.0 / n, am.data_ptr(),
                                       loss_sum.data_ptr(), _lib.stream_ptr()), 'sb_logits_ce')
    torch.cuda.synchronize()
    a32, b32 = dl_a.float(), dl_b.float()
    same = float((a32 == b32).float().mean())
    worst = float(((a32 - b32).abs() / (a32.abs() + 1e-12))[a32 != b32].max()) if same < 1 else 0.0
    print('B', B, 'identical gradient elements', same, 'worst relative difference of the others', worst)
    assert same > 0.995 and worst <= 2 ** -7 + 1e-6  # the others differ by one bf16 ulp (exp-sum order)
    for x, y, name in zip(ga, gb, ('d hf', 'd weight', 'd bias')):
        rel = float((x.float() - y.float()).abs().max() / (x.float().abs().max() + 1e-12))
        assert rel < 5e-3, (name, rel)
    # ---- rollout variant: arg-max only, nothing else written
    am2 = ops.logits_argmax(spec, hf, m.bias, pw)
    assert torch.equal(am2, amax_a)
