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
