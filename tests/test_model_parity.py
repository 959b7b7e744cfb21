"""End-to-end parity of the CUDA NextFramePredictor against the CPU oracle on identical weights, inputs and
injected randomness (north_star: bits bit-exact, logits within 1e-2 relative, loss / reward within 1e-3... of scale)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

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


@pytest.mark.parametrize('conditioned', [False, True])
def test_train_forward_backward_parity(conditioned):
    B, steps, eps = 2, 2, 0.3
    O, cfg, model, p0, frames_u8, per_step = _setup(B, steps, conditioned)
    from simple_b200.next_frame_predictor import InjectedRand
    from simple_b200.trainer import wm_loss
    ocfg = O.default_config(device='cpu', batch_size=B)
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    state = O.init_state(ocfg, B)
    frames = frames_u8.float() / 255
    model.train()
    model.init_internal_states(B)
    torch.manual_seed(123)
    np.random.seed(123)
    for j in range(steps):
        s = per_step[j]
        rand = O.GlobalRand()
        aux = {}
        tgt = s['new_state'].float() / 255
        logits, rew, val, lstm = O.forward(p, ocfg, state, frames, s['action'], tgt, eps, rand, training=True, aux=aux)
        loss, lrec, lval, lrew = O.wm_losses(ocfg, logits, rew, val, lstm, s['new_state'].long(), s['reward'], s['value'])
        names = list(p)
        gl = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
        ograd = dict(zip(names, gl))

        model.parity_rand = InjectedRand(rand.log, DEV)
        model.parity_rand.force = {'posterior_pre': aux['posterior_pre'].detach(), 'sampled_bits': aux['bits_pred']}
        model.zero_grad(set_to_none=True)
        out = wm_loss(model, cfg, frames.to(DEV), s['action'].to(DEV), s['new_state'].to(DEV), s['reward'].to(DEV),
                      s['value'].to(DEV), eps, keep_logits=True)
        out['loss'].backward()
        torch.cuda.synchronize()
        assert model.parity_rand.pos == len(rand.log), 'not all injected draws were consumed'
        # discrete latent decisions: natural end-to-end flip rate under bf16 upstream error (reported, bounded), then
        # forced to the oracle's decisions so that everything downstream is compared on identical bits
        mm = model.parity_rand.mismatch
        print('step', j, 'flip rates', mm, 'posterior_pre rel', _rel(model.last_aux['posterior_pre'], aux['posterior_pre']))
        assert mm['posterior_pre'] <= 0.03 and mm['sampled_bits'] <= 0.10, mm
        assert _rel(model.last_aux['posterior_pre'], aux['posterior_pre']) < 3e-2
        # `bits` mixes hard +-1 decisions with tanh values (mix, simple/utils.py:157): decisions exact, values to fp32 ulps
        assert torch.allclose(model.last_aux['bits'].detach().cpu(), aux['bits'].detach(), atol=1e-5), 'latent differs'
        # logits within 1e-2 relative (bf16 compute, fp32 accumulate)
        r = _rel(out['logits_ref'], logits)
        assert r < 1e-2, ('logits', j, r)
        assert _rel(out['reward_pred'], rew) < 1e-2
        # The value head reads the bottleneck right after an InstanceNorm(eps=1e-6) over only 3x2 = 6 pixels.  With
        # random-init weights ~6% of those (b,c) groups are ReLU-dead (variance < 1e-6, rstd ~ 1000), so a 1-ulp bf16
        # difference upstream moves a normalised value by O(1): the reference itself is ill-conditioned there
        # ("precision").  value_pred / loss_value therefore get a looser bound than the 1e-3 that holds for
        # the reconstruction, reward and lstm losses.
        assert _rel(out['value_pred'], val) < 1e-1
        for k, ref, tol in (('loss', loss, 1e-2), ('loss_reconstruct', lrec, 1e-3), ('loss_value', lval, 1e-1),
                            ('loss_reward', lrew, 1e-3), ('loss_lstm', lstm, 1e-3)):
            d = abs(float(out[k]) - float(ref))
            assert d < tol * max(1.0, abs(float(ref))), (k, j, float(out[k]), float(ref))
        # argmax decode: bit-exact given the same logits (stage-isolated), and report the end-to-end agreement
        am = out['argmax'].cpu().long()
        stage = torch.argmax(out['logits_ref'].cpu(), dim=1)
        assert torch.equal(am, stage), 'fused argmax differs from argmax of the same logits'
        # gradients.  Per-kernel backward parity is established in tests/test_ops.py on identical inputs (<= 2e-2).
        # End to end, bf16 vs fp32 differs at ReLU kinks (a ~0.5% flip rate of masks per layer costs ~sqrt(0.005) = 7%
        # relative L2 per block) and at the ill-conditioned 6-pixel InstanceNorms (see above), so the bound is tight
        # only for tensors downstream of those points and a direction (cosine) check elsewhere.
        import os
        os.makedirs('gpurun_out', exist_ok=True)
        rows = []
        for k, prm in model.named_parameters():
            og = ograd[k]
            if og is None:
                assert prm.grad is None or float(prm.grad.abs().max()) == 0, k
                continue
            assert prm.grad is not None, k
            gg = prm.grad.detach().float().cpu()
            rel_l2 = float((gg - og).norm() / (og.norm() + 1e-30))
            cos = float((gg * og).sum() / (gg.norm() * og.norm() + 1e-30))
            rows.append((k, rel_l2, cos, float(og.norm())))
        with open('gpurun_out/grad_errors_%s_step%d.txt' % ('conditioned' if conditioned else 'refinit', j), 'w') as f:
            for k, r, c, n in rows:
                f.write('%-60s relL2=%.4g cos=%.4f |g|=%.4g\n' % (k, r, c, n))
        tight = ('logits.', 'upscale_layers.9.', 'reward_estimator.', 'stochastic_model.bits_predictor.', 'value_estimator.')
        for k, r, c, n in rows:
            if k.startswith(tight):
                assert r < 0.15, (j, k, r)  # one flipped ReLU unit of the 128-wide reward MLP alone is ~9%
            if conditioned and n > 1e-9:  # (a bias in front of an all-active InstanceNorm has an exactly-zero gradient)
                assert c > 0.9, (j, k, r, c)
        onew = torch.argmax(logits, dim=1)
        frames = torch.cat((frames[:, 3:], onew.float() / 255), dim=1)
    model.parity_rand = None


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
        assert _rel(r2, rew) < 1e-2 and _rel(v2, val) < 8e-2
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
