"""Parity of the CUDA NextFramePredictor's reference-API forward against the pinned fp32 CPU oracle and the golden
fixtures recorded from the unmodified reference, on identical weights, inputs and injected randomness."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _setup(B, steps, conditioned=False):
    from oracle import wm_oracle as O
    from oracle.make_golden import make_inputs
    from simple_b200.next_frame_predictor import NextFramePredictor
    cfg = O.default_config(device=DEV, batch_size=B)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3)
    if conditioned:
        # Same architecture, but positive biases in front of the three InstanceNorms that normalise over <= 30 pixels:
        # no ReLU-dead (zero-variance) channels, so the network is well conditioned and gradients can be compared
        # tightly.  With the reference's random init those norms amplify any rounding by up to 1/sqrt(eps) = 1000.
        with torch.no_grad():
            for k in ('downscale_layers.6.bias', 'downscale_layers.8.bias', 'middle_network.middle_network.0.bias',
                      'middle_network.middle_network.2.bias', 'upscale_layers.0.bias'):
                dict(model.named_parameters())[k].add_(1.0)
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    frames_u8, per_step = make_inputs(1, B, 3, steps)
    return O, cfg, model, p0, frames_u8, per_step


# The end-to-end TRAINING comparison (forward, losses, gradients; unforced decisions; B = 16 and n_action = 9) lives in
# tests/test_strict_parity.py against the bf16-emulating oracle.  Here: the reference-API forward against the pinned fp32
# oracle and the recorded golden run of the unmodified reference.


def test_rollout_forward_parity_and_golden():
    B = 2
    O, cfg, model, p0, frames_u8, per_step = _setup(B, 2)
    from simple_b200.next_frame_predictor import InjectedRand
    golden = torch.load('tests/golden/wm_golden.pt')
    ocfg = O.default_config(device='cpu', batch_size=B)
    state = O.init_state(ocfg, B)
    frames = frames_u8.float() / 255
    model.eval()
    model.init_internal_states(B)
    torch.manual_seed(7)
    for j in range(2):
        rand = O.GlobalRand()
        aux = {}
        with torch.no_grad():
            logits, rew, val, _ = O.forward(p0, ocfg, state, frames, per_step[j]['action'].unsqueeze(1), None, 0, rand,
                                            training=False, aux=aux)
        g = golden['rollout'][j]
        assert torch.equal(aux['bits'], g['bits'])  # the oracle reproduces the reference's recorded run
        model.parity_rand = InjectedRand(rand.log, DEV)
        model.parity_rand.force = {'sampled_bits': aux['bits']}
        with torch.no_grad():
            lg, r2, v2 = model(frames.to(DEV), per_step[j]['action'].unsqueeze(1).to(DEV))
        torch.cuda.synchronize()
        print('rollout step', j, 'sampled-bit flip rate', model.parity_rand.mismatch)
        assert model.parity_rand.mismatch['sampled_bits'] <= 0.10
        assert _rel(lg, logits) < 1e-2, _rel(lg, logits)
        # value head on the reference init: 6-pixel InstanceNorm with eps = 1e-6 -- bf16 storage
        # plus the order of the split-K fp32 adds move it by up to ~0.08 of its range from run to run (strict-parity
        # tests bound it at 0.1 here and 2e-2 on the conditioned variant)
        assert _rel(r2, rew) < 1e-2 and _rel(v2, val) < 0.12
        new = torch.argmax(logits, dim=1)
        frames = torch.cat((frames[:, 3:], new.float() / 255), dim=1)
    model.parity_rand = None


def test_latent_sampling_bit_exact_stage_isolated():
    """Same bottleneck activations + same Exp(1) noise in => same sampled bits out (bit-exact), fp32 on both sides."""
    B = 4
    O, cfg, model, p0, _, _ = _setup(B, 1)
    from simple_b200.next_frame_predictor import InjectedRand
    ocfg = O.default_config(device='cpu', batch_size=B)
    torch.manual_seed(11)
    layer = torch.randn(B, 768, 3, 2)
    rand = O.GlobalRand()
    with torch.no_grad():
        obits, _ = O.bits_predictor(p0, ocfg, layer, rand)
    r = InjectedRand(rand.log, DEV)
    with torch.no_grad():
        bits, _ = model.stochastic_model.bits_predictor(layer.to(DEV).reshape(B, -1), r)
    assert torch.equal(bits.cpu(), obits)
