"""Host-side mirror of the reference surface on the GPU: Trainer.train on a synthetic replay list, the simulated env
(eager vs CUDA-graph replay, reference quirks), and checkpoint compatibility."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _cfg(**kw):
    from oracle.wm_oracle import default_config
    return default_config(device=DEV, **kw)


def _fake_buffer(n, n_action=3, seed=0):
    from bench import synth_replay
    d = synth_replay(n, n_action, seed)
    buf = []
    for i in range(n):
        buf.append([d['obs'][i], d['action'][i], d['reward'][i], d['new_obs'][i], torch.tensor(0, dtype=torch.uint8),
                    d['value'][i].clone()])
    return buf


def test_trainer_runs_on_reference_buffer_format_and_learns():
    from types import SimpleNamespace
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.trainer import Trainer
    cfg = _cfg(batch_size=4, rollout_length=6)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    trainer = Trainer(model, cfg)
    env = SimpleNamespace(buffer=_fake_buffer(40))
    m1 = trainer.train(1, env, steps=12)   # 2 windows: eager first steps, warm-up, capture, replay
    m2 = trainer.train(1, env, steps=36)
    assert set(m1) == {'loss', 'loss_reconstruct', 'loss_value', 'loss_reward', 'loss_lstm'}
    assert all(np.isfinite(v) for v in m2.values())
    assert m2['loss_reconstruct'] < m1['loss_reconstruct'], (m1, m2)  # the synthetic sprites are easy to fit
    gate_step = trainer.optimizer.state[model.gate.weight]['step']
    other_step = trainer.optimizer.state[model.logits.weight]['step']
    # gate.* has no gradient on the first step of every window (adafactor.py:125-126): its counter lags
    assert other_step == 48 and gate_step == 48 - 8, (gate_step, other_step)
    # the device-side step counters agree with the host-visible ones
    assert float(trainer.optimizer.state[model.gate.weight]['step_dev']) == gate_step
    env_bad = SimpleNamespace(buffer=[r[:5] + [None] for r in _fake_buffer(10)])
    with pytest.raises(BufferError):
        trainer.train(1, env_bad, steps=6)


def test_simulated_env_surface_and_graph_equivalence():
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.simulated_env import Discrete, make_simulated_env
    A = 4
    cfg = _cfg(agents=A, rollout_length=5)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    env = make_simulated_env(cfg, model, Discrete(3))
    assert env.num_envs == A and env.observation_space.shape == (12, 105, 80) and env.action_space.n == 3
    with pytest.raises(ValueError):
        env.reset()
    from bench import synth_replay
    init = synth_replay(A, 3, 5)['obs']
    for j in range(A):
        env.env_method('restart', init[j], indices=j)
    obs = env.reset()
    assert obs.dtype == torch.float32 and torch.equal(obs.cpu(), init.float())
    rewards = []
    for t in range(10):  # two rollouts of 5: eager first steps, warm-up, capture, replay
        obs, rew, done, infos = env.step(torch.randint(0, 3, (A, 1)))
        assert obs.shape == (A, 12, 105, 80) and rew.shape == (A, 1) and done.shape == (A,) and len(infos) == A
        assert not done.any()  # reference quirk: `done` is never True
        assert float(obs.min()) >= 0 and float(obs.max()) <= 255 and torch.equal(obs, obs.round())
        rewards.append(rew)
        if t == 4:
            assert env.step_count == 0  # wrapped: next step restarts from the initial frames
    r = torch.cat(rewards, 1).cpu()
    plain = torch.cat((r[:, :4], r[:, 5:9]), 1)
    assert set(plain.unique().tolist()) <= {-1.0, 0.0, 1.0}   # argmax - 1
    assert not set(r[:, 4].tolist()) <= {-1.0, 0.0, 1.0}      # last step of a rollout carries the value-head bonus
    # the first three channels of the new stack are the previous stack shifted by one frame
    env2 = make_simulated_env(cfg, model, Discrete(3))
    env2.restart_all(init.to(DEV))
    env2.reset()
    o1, *_ = env2.step(torch.zeros(A, 1, dtype=torch.long))
    assert torch.equal(o1[:, :9].cpu(), init[:, 3:].float())
    # batched restart from the real replay: valid records only, last agent = first small rollout
    replay = synth_replay(12, 3, 9)['obs']
    valid = torch.tensor([i % 3 != 0 for i in range(12)])
    first = torch.full((12, 105, 80), 7, dtype=torch.uint8)
    pick = env2.restart_from_replay(replay, valid=valid, first_small_rollout=first)
    assert bool(valid[pick.cpu()].all())
    got = env2.reset()
    assert torch.equal(got[-1].cpu(), first.float())
    assert torch.equal(got[:-1].cpu(), replay[pick.cpu()[:-1]].float())


def test_state_dict_round_trip_with_reference_keys():
    from simple_b200.next_frame_predictor import NextFramePredictor
    golden = torch.load('tests/golden/wm_golden.pt')
    cfg = _cfg()
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3)
    keys = list(model.state_dict().keys())
    assert keys == list(golden['train'][0]['param_sample'].keys())  # the reference's 120 keys, same order
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    m2 = NextFramePredictor(cfg, 3)
    m2.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))
    with pytest.raises(ValueError):
        NextFramePredictor(_cfg(hidden_size=64), 3)


def test_packed_gradient_mode_matches_reference_layout_mode():
    """Trainer fast path: conv-weight gradients stay in the packed GEMM-operand layout from wgrad to Adafactor, which
    also rewrites the bf16 operands.  Fed the SAME gradients (the packed buffers, unpacked for the reference-layout
    optimiser), both paths must produce the same weights, optimiser state and operands -- over two steps, the second
    one with the gate conv active.  (Two independent training runs are not comparable: fp32 atomics make the
    sign-like first Adafactor steps diverge run to run.)"""
    from bench import synth_replay
    from simple_b200.adafactor import Adafactor
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.ops import PackedWeight
    from simple_b200.trainer import wm_loss
    B = 2
    cfg = _cfg(batch_size=B)
    d = {k: v.to(DEV) for k, v in synth_replay(B + 4, 3, 3).items()}
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    torch.manual_seed(0)
    twin = NextFramePredictor(cfg, 3).to(DEV)  # same initial weights; only ever sees gradients copied from `model`
    opt, opt_ref = Adafactor(model.parameters()), Adafactor(twin.parameters())
    model.train()
    model.init_internal_states(B)
    frames = d['obs'][:B].float() / 255
    for j in range(2):
        idx = torch.arange(B, device=DEV) + j
        out = wm_loss(model, cfg, frames, d['action'][idx].float(), d['new_obs'][idx], d['reward'][idx].long(),
                      d['value'][idx], 0)
        opt.zero_grad(set_to_none=True)
        packed = model.packed_weight_map(defer_unpack=True)
        out['loss'].backward()
        by_name = {k: pw for k, pw in model._packed.items()}
        mods, tmods = model._weights(), twin._weights()
        conv_w = {id(mods[k].weight): k for k in mods}
        for (k, p), q in zip(model.named_parameters(), twin.parameters()):
            if id(p) in conv_w:
                pw = by_name[conv_w[id(p)]]
                assert p.grad is None  # never materialised in the reference layout
                q.grad = pw.unpack_grad(pw.gpacked) if pw.gpacked is not None else None
            else:
                q.grad = p.grad.clone() if p.grad is not None else None
        if j == 0:
            assert by_name['gate'].gpacked is None and twin.gate.weight.grad is None
        else:
            assert by_name['gate'].gpacked is not None
        opt.step(max_grad_norm=1.0, packed=packed)
        opt_ref.step(max_grad_norm=1.0)
        model.commit_state()
        torch.cuda.synchronize()
        assert abs(float(opt.grad_norm_sq) - float(opt_ref.grad_norm_sq)) < 1e-5 * float(opt_ref.grad_norm_sq)
        for (k, a), b in zip(model.named_parameters(), twin.parameters()):
            err = float((a - b).abs().max() / (a.abs().max() + 1e-12))
            assert err < 1e-5, (j, k, err)
            sa, sb_ = opt.state[a], opt_ref.state[b]
            assert sa.get('step', 0) == sb_.get('step', 0)
            for key in ('exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq'):
                if key in sa:
                    assert torch.allclose(sa[key], sb_[key], rtol=1e-4, atol=1e-30), (j, k, key)
        # operands written by the optimiser == operands packed from the (bitwise own) master weights
        for k, pw in model._packed.items():
            if pw.gpacked is None:
                continue
            assert pw.fresh
            ref = PackedWeight(mods[k].weight, pw.s, Ipad=pw.Ipad, Opad=pw.Opad,
                               iperm=pw.iperm.tolist() if pw.iperm is not None else None,
                               operm=pw.operm.tolist() if pw.operm is not None else None)
            assert torch.equal(ref.fwd, pw.fwd) and torch.equal(ref.tr, pw.tr), (j, k)
        frames = torch.cat((frames[:, 3:], out['argmax'].float() / 255), dim=1)


def test_checkpoint_resume_restores_weights_optimizer_state_and_rng():
    """a checkpoint carries weights, Adafactor state (incl. the lagging gate.* counters), dropout seed
    and RNG streams; a fresh Trainer that loads it continues with the same optimiser state."""
    import os
    import tempfile
    from types import SimpleNamespace
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.trainer import Trainer
    cfg = _cfg(batch_size=2, rollout_length=4)
    cfg.cuda_graph = False
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    trainer = Trainer(model, cfg)
    env = SimpleNamespace(buffer=_fake_buffer(24))
    trainer.train(1, env, steps=8)  # two windows of 4 steps
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'ck.pt')
        trainer.save_checkpoint(path, epoch=3)
        torch.manual_seed(123)
        model2 = NextFramePredictor(cfg, 3).to(DEV)
        trainer2 = Trainer(model2, cfg)
        assert trainer2.load_checkpoint(path) == 3
    for (k, a), b in zip(model.named_parameters(), model2.parameters()):
        assert torch.equal(a, b), k
        sa, sb_ = trainer.optimizer.state[a], trainer2.optimizer.state[b]
        assert sa['step'] == sb_['step'] and float(sb_['step_dev']) == sb_['step'], k
        for key in ('exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq'):
            if key in sa:
                assert torch.equal(sa[key], sb_[key]), (k, key)
    assert trainer.optimizer.state[model.gate.weight]['step'] == 8 - 2   # skipped on the first step of each window
    assert int(model2._seed_dev) == int(model._seed_dev)
    assert torch.equal(torch.cuda.get_rng_state(torch.device(DEV)), torch.cuda.get_rng_state(torch.device(DEV)))
    m = trainer2.train(1, env, steps=4)  # and it keeps training
    assert all(np.isfinite(v) for v in m.values())
    assert trainer2.optimizer.state[model2.logits.weight]['step'] == 12


def test_device_ppo_learns_in_the_simulated_env():
    """PPO.learn on the device-resident simulated env (rollout storage, GAE, clipped-surrogate update)
    runs end to end without host round trips and changes the policy."""
    from bench import synth_replay
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.policy import CnnPolicy
    from simple_b200.ppo import DevicePPO
    from simple_b200.simulated_env import Discrete, make_simulated_env
    A = 4
    cfg = _cfg(agents=A, rollout_length=6)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    env = make_simulated_env(cfg, model, Discrete(3))
    env.restart_all(synth_replay(A, 3, 5)['obs'].to(DEV))
    policy = CnnPolicy((12, 105, 80), 3).to(DEV)
    before = [p.detach().clone() for p in policy.parameters()]
    ppo = DevicePPO(env, policy, num_mini_batch=2, ppo_epoch=2)
    st = ppo.collect()
    assert st['obs'].dtype == torch.uint8 and st['obs'].shape == (7, A, 12, 105, 80)
    assert st['returns'].shape == (7, A, 1) and bool(torch.isfinite(st['returns'][:-1]).all())
    # GAE against the reference recursion on the host
    v, r, m = st['values'].cpu(), st['rewards'].cpu(), st['masks'].cpu()
    gae, want = torch.zeros(A, 1), torch.zeros(6, A, 1)
    for t in reversed(range(6)):
        delta = r[t] + 0.99 * v[t + 1] * m[t + 1] - v[t]
        gae = delta + 0.99 * 0.95 * m[t + 1] * gae
        want[t] = gae + v[t]
    assert torch.allclose(st['returns'][:-1].cpu(), want, atol=1e-5)
    metrics = ppo.learn(2 * 6 * A)
    assert set(metrics) == {'ppo_value_loss', 'ppo_action_loss', 'ppo_entropy'}
    assert all(np.isfinite(x) for x in metrics.values()) and 0 < metrics['ppo_entropy'] <= np.log(3) + 1e-4
    assert any(not torch.equal(a, b) for a, b in zip(before, policy.parameters()))


@pytest.mark.gpu
def test_env_post_kernel_matches_torch_ops():
    """sb_env_post (frame-stack shift in place + float observation + reward decode + value copy) vs the torch ops of
    simple/subproc_vec_env.py:428-441 it replaces."""
    from simple_b200 import _lib
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(3)
    for A, SC, c, H, W in ((16, 12, 3, 105, 80), (3, 8, 2, 4, 8)):
        frames = torch.randint(0, 256, (A, SC, H, W), generator=g, dtype=torch.uint8).to(dev)
        fresh = torch.randint(0, 256, (A, c, H, W), generator=g, dtype=torch.uint8).to(dev)
        rp = torch.randn(A, 3, generator=g).to(dev)
        rp[0] = 0.5  # a three-way tie decodes to the first maximum
        vp = torch.randn(A, generator=g).to(dev)
        want_frames = torch.cat((frames[:, c:], fresh), dim=1)
        obs = torch.full((A, SC, H, W), -1.0, device=dev)
        rew, val = torch.empty(A, device=dev), torch.empty(A, device=dev)
        _lib.check(_lib.lib().sb_env_post(frames.data_ptr(), fresh.data_ptr(), obs.data_ptr(), A, SC, c, H * W,
                                          rp.data_ptr(), 3, vp.data_ptr(), rew.data_ptr(), val.data_ptr(),
                                          _lib.stream_ptr()), 'sb_env_post')
        torch.cuda.synchronize()
        assert torch.equal(frames, want_frames)
        assert torch.equal(obs, want_frames.float())
        assert torch.equal(rew, (torch.argmax(rp, dim=1) - 1).float()) and float(rew[0]) == -1.0
        assert torch.equal(val, vp)
    # unsupported geometry is refused, not mis-handled
    assert _lib.lib().sb_env_post(frames.data_ptr(), fresh.data_ptr(), None, 1, 8, 2, 24, None, 0, None, None, None,
                                  _lib.stream_ptr()) != 0
