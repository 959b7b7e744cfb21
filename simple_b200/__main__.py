"""`python -m simple_b200`: the flag surface of the reference's `python -m simple` (simple/__main__.py:153-192).

The epoch orchestration of the reference (real-Atari collection, PPO update, evaluation) is OUT OF SCOPE of this
B200 hot-path build (needs gym/ALE/baselines).  This entry point parses the same flags, builds the
world model + trainer + simulated env on the GPU and runs a self-contained demonstration on a synthetic replay:
world-model training followed by PPO iterations in the simulated env (device-resident rollouts + updates).
"""
import argparse
from time import strftime

import torch


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--agents', type=int, default=16)
    p.add_argument('--batch-size', type=int, default=4)
    p.add_argument('--bottleneck-bits', type=int, default=128)
    p.add_argument('--bottleneck-noise', type=float, default=0.1)
    p.add_argument('--clip-grad-norm', type=float, default=1.0)
    p.add_argument('--compress-steps', type=int, default=5)
    p.add_argument('--device', type=str, default='cuda')
    p.add_argument('--done-on-last-rollout-step', default=True, action='store_false')
    p.add_argument('--dropout', type=float, default=0.15)
    p.add_argument('--env-name', type=str, default='Freeway')
    p.add_argument('--epochs', type=int, default=15)
    p.add_argument('--experiment-name', type=str, default=strftime('%d-%m-%y-%H:%M:%S'))
    p.add_argument('--filter-double-steps', type=int, default=3)
    p.add_argument('--frame-shape', type=int, nargs=3, default=(3, 105, 80))
    p.add_argument('--hidden-layers', type=int, default=2)
    p.add_argument('--hidden-size', type=int, default=96)
    p.add_argument('--input-noise', type=float, default=0.05)
    p.add_argument('--latent-rnn-max-sampling', type=float, default=0.5)
    p.add_argument('--latent-state-size', type=int, default=128)
    p.add_argument('--latent-use-max-probability', type=float, default=0.8)
    p.add_argument('--load-models', default=False, action='store_true')
    p.add_argument('--noop-max', type=int, default=8)
    p.add_argument('--ppo-gamma', type=float, default=0.99)
    p.add_argument('--ppo-lr', type=float, default=1e-4)
    p.add_argument('--recurrent-state-size', type=int, default=64)
    p.add_argument('--render-evaluation', default=False, action='store_true')
    p.add_argument('--render-training', default=False, action='store_true')
    p.add_argument('--residual-dropout', type=float, default=0.5)
    p.add_argument('--rollout-length', type=int, default=50)
    p.add_argument('--save-models', default=False, action='store_true')
    p.add_argument('--scheduled-sampling-decay-steps', type=int, default=22250)
    p.add_argument('--simulation-flip-first-random-for-beginning', default=True, action='store_false')
    p.add_argument('--stacking', type=float, default=4)
    p.add_argument('--stack-internal-states', default=True, action='store_false')
    p.add_argument('--target-loss-clipping', type=float, default=0.03)
    p.add_argument('--use-ppo-lr-decay', default=False, action='store_true')
    p.add_argument('--use-stochastic-model', default=True, action='store_false')
    p.add_argument('--use-wandb', default=False, action='store_true')
    # demo-only
    p.add_argument('--wm-steps', type=int, default=100, help='world-model optimiser steps on the synthetic replay')
    p.add_argument('--n-action', type=int, default=3)
    p.add_argument('--ppo-iterations', type=int, default=2, help='PPO.learn iterations in the simulated env')
    return p


def main():
    config = build_parser().parse_args()
    from types import SimpleNamespace
    import os
    from .synthetic import synth_replay
    from .next_frame_predictor import NextFramePredictor
    from .policy import CnnPolicy
    from .simulated_env import Discrete, make_simulated_env
    from .trainer import Trainer
    dev = torch.device(config.device if config.device != 'cuda' else 'cuda:0')
    config.device = str(dev)
    model = NextFramePredictor(config, config.n_action).to(dev)
    if config.load_models:
        model.load_state_dict(torch.load(os.path.join('models', 'model.pt')))
    trainer = Trainer(model, config)
    n = config.rollout_length + max(64, config.batch_size) + 8
    d = synth_replay(n, config.n_action, seed=0)
    buf = [[d['obs'][i], d['action'][i], d['reward'][i], d['new_obs'][i], torch.tensor(0, dtype=torch.uint8),
            d['value'][i].clone()] for i in range(n)]
    if config.save_models and not os.path.isdir('models'):
        os.mkdir('models')
    metrics = trainer.train(1, SimpleNamespace(buffer=buf), steps=config.wm_steps)
    print('world model:', metrics)
    env = make_simulated_env(config, model, Discrete(config.n_action))
    policy = CnnPolicy(env.observation_space.shape, config.n_action).to(dev)
    from .ppo import DevicePPO
    ppo = DevicePPO(env, policy, gamma=config.ppo_gamma, lr=config.ppo_lr)
    for it in range(config.ppo_iterations):
        # restart path of simple/__main__.py:85-91: every agent from a random replay record, the last one from the
        # first frames of the (synthetic) episode
        env.restart_from_replay(d['obs'], first_small_rollout=d['obs'][0],
                                flip_first=config.simulation_flip_first_random_for_beginning)
        print(f'ppo iteration {it}:', ppo.learn(config.rollout_length * config.agents))


if __name__ == '__main__':
    main()
