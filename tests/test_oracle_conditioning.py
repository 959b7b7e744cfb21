"""Evidence for DESIGN.md section 5 ("precision"): on the REFERENCE init the world model's training gradient is an
ill-conditioned -- in places discontinuous -- function of its own forward pass.  The InstanceNorms over 3x2 and 6x5
pixels use eps = 1e-6, i.e. a gain of up to 1000 on ReLU-sparse channels, and the ReLU kinks in front of them switch
that gain on and off.  Perturbing the deep conv outputs of the ORACLE by 1e-6 relative (what a different fp32 summation
order does) moves the oracle's own encoder gradient by percent to hundreds of percent, while everything downstream of
the bottleneck barely moves; with positive biases in front of those norms (the "conditioned" variant used by the
gradient-parity tests) the same perturbation is harmless.  Pure CPU, no CUDA."""
import numpy as np
import torch
import torch.nn.functional as F


def _grads(p0, emulate, noise, seed):
    from oracle import wm_oracle as O
    from oracle.make_golden import make_inputs
    B = 2
    frames_u8, per_step = make_inputs(1, B, 3, 1)
    g = torch.Generator().manual_seed(seed)
    real = F.conv2d

    def conv(x, w, b=None, **kw):
        y = real(x, w, b, **kw)
        if noise > 0 and y.shape[2] * y.shape[3] <= 130:
            y = y * (1 + noise * torch.randn(y.shape, generator=g))
        return y
    O.F.conv2d = conv
    try:
        c = O.default_config(device='cpu', batch_size=B, emulate_bf16=emulate)
        p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
        torch.manual_seed(123)
        np.random.seed(123)
        s = per_step[0]
        logits, rew, val, lstm = O.forward(p, c, O.init_state(c, B), frames_u8.float() / 255, s['action'],
                                           s['new_state'].float() / 255, 0.3, O.GlobalRand(), training=True)
        loss, *_ = O.wm_losses(c, logits, rew, val, lstm, s['new_state'].long(), s['reward'], s['value'])
        return dict(zip(p, torch.autograd.grad(loss, list(p.values()), allow_unused=True)))
    finally:
        O.F.conv2d = real


def _init(conditioned):
    from oracle import wm_oracle as O
    from simple_b200.next_frame_predictor import NextFramePredictor
    torch.manual_seed(0)
    m = NextFramePredictor(O.default_config(device='cpu', batch_size=2), 3)
    if conditioned:
        with torch.no_grad():
            for k in ('downscale_layers.6.bias', 'downscale_layers.8.bias', 'middle_network.middle_network.0.bias',
                      'middle_network.middle_network.2.bias', 'upscale_layers.0.bias'):
                dict(m.named_parameters())[k].add_(1.0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _rel(a, b, k):
    return float((a[k] - b[k]).norm() / (b[k].norm() + 1e-30))


def test_training_gradient_is_ill_conditioned_in_the_reference_itself():
    """1e-6 relative noise on the deep conv outputs (= another fp32 summation order):
         fp32 oracle, reference init:   encoder gradient moves ~1e-4 (gain ~100), decoder ~5e-7
         fp32 oracle, conditioned:      encoder ~5e-6
         bf16-emulating oracle:         the noise flips bf16 roundings (0.4 % jumps) that the 6-pixel InstanceNorms
                                        amplify: encoder gradient moves by several PERCENT, reference init and conditioned
       => encoder-gradient parity of any bf16 implementation against any oracle is bounded by this, not by the kernels."""
    enc, dec = 'downscale_layers.0.weight', 'upscale_layers.8.weight'
    p_ref, p_cond = _init(False), _init(True)
    rows = {}
    for name, p0, emul in (('fp32/reference-init', p_ref, False), ('fp32/conditioned', p_cond, False),
                           ('bf16/reference-init', p_ref, True)):
        base = _grads(p0, emul, 0.0, 0)
        pert = [_grads(p0, emul, 1e-6, s) for s in (1, 2)]
        rows[name] = (max(_rel(g, base, enc) for g in pert), max(_rel(g, base, dec) for g in pert))
        print(name, 'encoder grad moves %.2e, decoder %.2e' % rows[name])
    assert rows['fp32/reference-init'][0] > 10 * rows['fp32/conditioned'][0]   # the reference init is the ill-conditioned one
    assert rows['fp32/reference-init'][0] > 20 * rows['fp32/reference-init'][1]  # and only upstream of the bottleneck
    assert rows['bf16/reference-init'][0] > 1e-2                                  # bf16 storage: percent-level chaos
