"""Pin the oracle against the live reference and write tests/golden/*.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

1. Builds the reference NextFramePredictor (n_action=3, default flags) under torch.manual_seed(0).
2. Runs 3 training steps of the reference (forward, trainer.py losses, backward, clip, Adafactor) and
   2 eval/rollout forwards, with torch/numpy seeded; runs the oracle restatement on the same weights
   with `GlobalRand` from the same seeds; asserts they agree (bit-exact where integer, 1e-5 otherwise).
3. Stores compact fixtures (scalars, checksums, strided samples) in tests/golden/wm_golden.pt with
   the seeds needed to regenerate the inputs and weights deterministically anywhere.
"""
import copy
import os
import sys

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)

from oracle import wm_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def make_inputs(seed, batch, n_action, steps):
    """Deterministic synthetic 'Atari-like' inputs: u8 frames, one-hot actions, rewards, values."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (batch, 12, 105, 80), generator=g, dtype=torch.uint8)
    # make it less uniform: a constant background on half of the image
    frames[:, :, 50:, :] = 17
    per_step = []
    for _ in range(steps):
        a = torch.randint(0, n_action, (batch,), generator=g)
        per_step.append(dict(
            action=torch.nn.functional.one_hot(a, n_action).float(),
            new_state=torch.randint(0, 256, (batch, 3, 105, 80), generator=g, dtype=torch.uint8),
            reward=torch.randint(0, 3, (batch,), generator=g),
            value=torch.randn(batch, generator=g),
        ))
    return frames, per_step


def build_reference_model(ref, cfg, n_action, seed):
    torch.manual_seed(seed)
    return ref.nfp.NextFramePredictor(cfg, n_action)


def sample(t, n=64):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return f[idx].float().clone()


def main():
    ref = load_reference()
    assert ref is not None, 'reference not reachable'
    torch.set_num_threads(os.cpu_count())
    cfg = O.default_config(device='cpu', batch_size=2)
    n_action, B, steps = 3, 2, 3
    model = build_reference_model(ref, cfg, n_action, seed=0)
    frames_u8, per_step = make_inputs(1, B, n_action, steps)

    # ---------------- reference: 3 training steps ----------------
    opt = ref.adafactor.Adafactor(model.parameters())
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.manual_seed(123)
    np.random.seed(123)
    model.train()
    model.init_internal_states(B)
    frames = frames_u8.float() / 255
    eps = 0.3
    ref_out = []
    for j in range(steps):
        s = per_step[j]
        tgt_in = s['new_state'].float() / 255
        logits, rew, val = model(frames, s['action'], tgt_in, eps)
        lstm = model.stochastic_model.get_lstm_loss()
        loss, lrec, lval, lrew = O.wm_losses(cfg, logits, rew, val, lstm, s['new_state'].long(), s['reward'], s['value'])
        opt.zero_grad()
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.clip_grad_norm)
        grads = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in model.named_parameters()}
        opt.step()
        new = torch.argmax(logits, dim=1)
        ref_out.append(dict(logits=logits.detach(), reward=rew.detach(), value=val.detach(), loss=loss.detach(),
                            lrec=lrec.detach(), lval=lval.detach(), lrew=lrew.detach(), lstm=lstm.detach(),
                            gnorm=gnorm.detach(), grads=grads, argmax=new,
                            params={k: v.detach().clone() for k, v in model.state_dict().items()}))
        frames = torch.cat((frames[:, 3:], new.float() / 255), dim=1)

    # ---------------- oracle: same thing ----------------
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    torch.manual_seed(123)
    np.random.seed(123)
    rand = O.GlobalRand()
    state = O.init_state(cfg, B)
    frames = frames_u8.float() / 255
    opt_state = {k: {} for k in p}
    worst = worst_g = worst_p = 0.0
    golden = dict(seed_model=0, seed_inputs=1, seed_rng=123, batch=B, n_action=n_action, steps=steps, eps=eps, train=[])
    for j in range(steps):
        s = per_step[j]
        tgt_in = s['new_state'].float() / 255
        logits, rew, val, lstm = O.forward(p, cfg, state, frames, s['action'], tgt_in, eps, rand, training=True)
        loss, lrec, lval, lrew = O.wm_losses(cfg, logits, rew, val, lstm, s['new_state'].long(), s['reward'], s['value'])
        names = [k for k in p]
        gl = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
        present = [(k, g) for k, g in zip(names, gl) if g is not None]
        clipped, gnorm = O.clip_grad_norm([g for _, g in present], cfg.clip_grad_norm)
        r = ref_out[j]
        new = torch.argmax(logits, dim=1)
        assert torch.equal(new, r['argmax']), 'argmax frames differ'
        for a, b, nm in ((logits, r['logits'], 'logits'), (rew, r['reward'], 'reward'), (val, r['value'], 'value'),
                         (loss, r['loss'], 'loss'), (lstm, r['lstm'], 'lstm'), (gnorm, r['gnorm'], 'gnorm')):
            d = float((a.detach() - b).abs().max() / (b.abs().max() + 1e-12))
            worst = max(worst, d)
            assert d < 1e-5, (nm, j, d)
        with torch.no_grad():
            for (k, _), g in zip(present, clipped):
                rg = r['grads'][k]
                assert rg is not None, k
                d = float((g - rg).abs().max() / (rg.abs().max() + 1e-20))
                worst_g = max(worst_g, d)
                assert d < 2e-3, ('grad', k, j, d)
                # stage-isolated optimiser check: the oracle's Adafactor is fed the REFERENCE's clipped gradient
                # (at step 1 the update is sign(g), so noise-level gradients would otherwise flip elements)
                p[k].copy_(O.adafactor_step(p[k].detach(), rg, opt_state[k]))
            for k in names:
                if k not in dict(present):
                    assert r['grads'][k] is None, ('reference has grad where oracle has none', k)
                d = float((p[k].detach() - r['params'][k]).abs().max() / (r['params'][k].abs().max() + 1e-20))
                worst_p = max(worst_p, d)
                assert d < 1e-5, ('param', k, j, d)
                p[k].copy_(r['params'][k])  # keep both trajectories on identical weights
        golden['train'].append(dict(
            loss=float(r['loss']), lrec=float(r['lrec']), lval=float(r['lval']), lrew=float(r['lrew']),
            lstm=float(r['lstm']), gnorm=float(r['gnorm']), reward=r['reward'].clone(), value=r['value'].clone(),
            logits_sample=sample(r['logits'], 256), argmax_sum=int(r['argmax'].sum()),
            argmax_sample=sample(r['argmax'].float(), 256).long(),
            none_grads=sorted(k for k, v in r['grads'].items() if v is None),
            param_sample={k: sample(v, 8) for k, v in r['params'].items()},
            grad_sample={k: sample(v, 8) for k, v in r['grads'].items() if v is not None},
        ))
        frames = torch.cat((frames[:, 3:], new.float() / 255), dim=1)
    print('train parity oracle vs reference: worst rel diff fwd %.3g grad %.3g param %.3g' % (worst, worst_g, worst_p))

    # ---------------- eval / rollout forwards (reference subproc_vec_env.py:409-435 arithmetic) ----------------
    model.load_state_dict(p0)
    model.eval()
    torch.manual_seed(7)
    model.init_internal_states(B)
    frames = frames_u8.float() / 255
    golden['rollout'] = []
    ref_roll = []
    for j in range(2):
        with torch.no_grad():
            logits, rew, val = model(frames, per_step[j]['action'].unsqueeze(1))
        new = torch.argmax(logits, dim=1)
        frames = torch.cat((frames[:, 3:], new.float() / 255), dim=1)
        ref_roll.append((logits, rew, val, new, (frames * 255).byte()))
    pe = {k: v.clone() for k, v in p0.items()}
    torch.manual_seed(7)
    rand = O.GlobalRand()
    state = O.init_state(cfg, B)
    frames = frames_u8.float() / 255
    for j in range(2):
        aux = {}
        with torch.no_grad():
            logits, rew, val, _ = O.forward(pe, cfg, state, frames, per_step[j]['action'].unsqueeze(1), None, 0, rand,
                                            training=False, aux=aux)
        new = torch.argmax(logits, dim=1)
        frames = torch.cat((frames[:, 3:], new.float() / 255), dim=1)
        rl = ref_roll[j]
        assert torch.equal(new, rl[3]), 'rollout argmax differs'
        assert torch.equal((frames * 255).byte(), rl[4])
        assert float((logits - rl[0]).abs().max()) < 1e-4
        golden['rollout'].append(dict(argmax_sum=int(new.sum()), argmax_sample=sample(new.float(), 256).long(),
                                      reward=rl[1].clone(), value=rl[2].clone(), bits=aux['bits'].clone(),
                                      logits_sample=sample(rl[0], 256)))
    print('rollout parity oracle vs reference: ok')

    # ---------------- small known-answer vectors for helper functions ----------------
    torch.manual_seed(5)
    x = torch.rand(2, 3, 105, 80)
    golden['standardize'] = dict(seed=5, out_sample=sample(torch.stack([ref.utils.standardize_frame(f) for f in x]), 64))
    golden['timing'] = {str(s): sample(ref.utils.get_timing_signal_nd(s), 64) for s in ((96, 105, 80), (768, 6, 5))}
    assert torch.equal(ref.utils.get_timing_signal_nd((96, 105, 80)), O.timing_signal_nd((96, 105, 80)))
    bits = (torch.rand(4, 16, 8) > 0.5).float()
    assert torch.equal(ref.utils.bit_to_int(bits, 8).long(), O.bit_to_int(bits))
    ints = torch.randint(0, 256, (4, 16))
    assert torch.equal(ref.utils.int_to_bit(ints, 8), O.int_to_bit(ints))
    golden['bit_to_int'] = dict(bits=bits, ints=O.bit_to_int(bits))
    # MeanAttention known answer
    torch.manual_seed(9)
    ma = ref.utils.MeanAttention(16, 6)
    xin = torch.randn(2, 16, 6, 5)
    golden['mean_attention'] = dict(state={k: v.clone() for k, v in ma.state_dict().items()}, x=xin, out=ma(xin).detach())
    pm = {'m.' + k: v for k, v in ma.state_dict().items()}
    assert float((O.mean_attention(pm, 'm', xin) - golden['mean_attention']['out']).abs().max()) < 1e-6
    # torch.median (lower median) + input noise
    st = torch.randint(0, 256, (2, 12, 10, 8), dtype=torch.uint8)
    km = (torch.rand(2, 12, 10, 8) < 0.95).float()
    golden['input_noise'] = dict(state=st, keep=km, out=O.input_noise(st, km))

    out = os.path.join(_REPO, 'tests', 'golden', 'wm_golden.pt')
    torch.save(golden, out)
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
