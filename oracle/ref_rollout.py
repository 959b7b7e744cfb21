"""CPU baseline of the simulated rollout: the UNMODIFIED reference stack -- `simple.subproc_vec_env.make_simulated_env`
(16 echo subprocesses behind pipes) stepped by the reference `PPO` policy (`a2c_ppo_acktr.policy.Policy.act`) around the
reference NextFramePredictor on `device='cpu'`.  TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's rollout cpu_baseline).

    python -m oracle.ref_rollout --agents 16 --steps 8 --n-action 3      ->  one JSON line on stdout

The driver sits under `if __name__ == '__main__'` because the reference's SubprocVecEnv uses the forkserver start method
(simple/subproc_vec_env.py:274-280); the gym / baselines shims of oracle/shims are put on PYTHONPATH by ref_loader.
"""
import argparse
import json
import os
import sys
import time

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=16)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--n-action', type=int, default=3)
    args = ap.parse_args()
    import torch
    from oracle import wm_oracle as O
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is None:
        print(json.dumps({'unavailable': 'reference not reachable (baseline/_ref missing)'}))
        return
    torch.set_num_threads(os.cpu_count())
    import gym
    from simple.subproc_vec_env import make_simulated_env
    from atari_utils.ppo_wrapper import PPO
    from simple_b200.synthetic import synth_replay
    cfg = O.default_config(device='cpu', agents=args.agents, rollout_length=args.steps + args.warmup)
    torch.manual_seed(0)
    model = ref.nfp.NextFramePredictor(cfg, args.n_action)
    env = make_simulated_env(cfg, model, gym.spaces.Discrete(args.n_action))
    agent = PPO(env, cfg.device, gamma=cfg.ppo_gamma, num_steps=cfg.rollout_length, num_mini_batch=5, lr=cfg.ppo_lr)
    init = synth_replay(args.agents, args.n_action, 3)['obs']
    for j in range(args.agents):
        env.env_method('restart', init[j], indices=j)
    obs = env.reset()
    policy = agent.actor_critic
    t0 = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t0 = time.perf_counter()
        with torch.no_grad():
            _, action, _ = policy.act(obs)  # atari_utils/ppo_wrapper.py:102
        obs, reward, done, infos = env.step(action)
    dt = time.perf_counter() - t0
    env.close()
    print(json.dumps({'env_steps_per_s': args.agents * args.steps / dt, 'agents': args.agents, 'steps': args.steps,
                      'seconds': dt, 'cores': os.cpu_count(), 'n_action': args.n_action}))


if __name__ == '__main__':
    main()
