"""Per-kernel time table of a few training steps / rollout steps with torch.profiler (CUPTI).  Usage on the GPU box:
   python tools/profile_step.py [train|rollout] [batch]"""
import os
import sys
import torch
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_replay  # noqa: E402
from simple_b200.next_frame_predictor import NextFramePredictor  # noqa: E402
from simple_b200.trainer import Trainer  # noqa: E402
from simple_b200.simulated_env import Discrete, make_simulated_env  # noqa: E402
from simple_b200.policy import CnnPolicy  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device('cuda:0')
cfg = SimpleNamespace(
    agents=B, batch_size=B, bottleneck_bits=128, bottleneck_noise=0.1, clip_grad_norm=1.0, compress_steps=5,
    device=str(dev), done_on_last_rollout_step=True, dropout=0.15, filter_double_steps=3, frame_shape=(3, 105, 80),
    hidden_layers=2, hidden_size=96, input_noise=0.05, latent_rnn_max_sampling=0.5, latent_state_size=128,
    latent_use_max_probability=0.8, recurrent_state_size=64, residual_dropout=0.5, rollout_length=50,
    save_models=False, scheduled_sampling_decay_steps=22250, stacking=4, stack_internal_states=True,
    target_loss_clipping=0.03, use_stochastic_model=True, use_wandb=False)
torch.manual_seed(0)
model = NextFramePredictor(cfg, 3).to(dev)
data = {k: v.to(dev) for k, v in synth_replay(B + 20, 3, 1).items()}
starts = torch.arange(B, device=dev)
if mode == 'train':
    trainer = Trainer(model, cfg)
    frames = trainer.preprocess_state(data['obs'][starts], per_sample=False)
    model.init_internal_states(B)

    def step(j, frames):
        idx = starts + j
        out = trainer.train_step(frames, data['action'][idx].float(), data['new_obs'][idx], data['reward'][idx].long(),
                                 data['value'][idx], 0)
        nxt = trainer.preprocess_state(out['argmax'], per_sample=True)
        return torch.cat((frames[:, 3:], nxt), dim=1)
    for j in range(3):
        frames = step(j, frames)
    n = 3
else:
    env = make_simulated_env(cfg, model, Discrete(3))
    policy = CnnPolicy((12, 105, 80), 3).to(dev)
    env.restart_all(data['obs'][:B])
    obs = env.reset()

    def step(j, obs):
        _, a, _ = policy.act(obs)
        return env.step(a)[0]
    for j in range(3):
        obs = step(j, obs)
    frames = obs
    n = 5
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for j in range(3, 3 + n):
        frames = step(j, frames)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
tab = prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70)
os.makedirs('gpurun_out', exist_ok=True)
with open(f'gpurun_out/profile_{mode}_b{B}.txt', 'w') as f:
    f.write(f'{mode} B={B}: {n} steps, wall {wall * 1e3 / n:.2f} ms/step (under profiler)\n')
    f.write(tab)
print(tab[:200])
