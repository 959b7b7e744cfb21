"""A/B of the split-K rounding rule on the graph-replayed training step (ms per step, CUDA events).
python tools/split_ab.py"""
import os
import sys
import torch
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_replay  # noqa: E402
from simple_b200 import ops  # noqa: E402
from simple_b200.next_frame_predictor import NextFramePredictor  # noqa: E402
from simple_b200.trainer import StepRunner, Trainer  # noqa: E402

B = 64
dev = torch.device('cuda:0')
cfg = SimpleNamespace(
    agents=16, batch_size=B, bottleneck_bits=128, bottleneck_noise=0.1, clip_grad_norm=1.0, compress_steps=5,
    device=str(dev), done_on_last_rollout_step=True, dropout=0.15, filter_double_steps=3, frame_shape=(3, 105, 80),
    hidden_layers=2, hidden_size=96, input_noise=0.05, latent_rnn_max_sampling=0.5, latent_state_size=128,
    latent_use_max_probability=0.8, recurrent_state_size=64, residual_dropout=0.5, rollout_length=50,
    save_models=False, scheduled_sampling_decay_steps=22250, stacking=4, stack_internal_states=True,
    target_loss_clipping=0.03, use_stochastic_model=True, use_wandb=False)
torch.manual_seed(0)
model = NextFramePredictor(cfg, 3).to(dev)
trainer = Trainer(model, cfg)
d = {k: v.to(dev) for k, v in synth_replay(B + 60, 3, 1).items()}
replay = dict(obs=d['obs'], new_obs=d['new_obs'], action=d['action'].float(), reward=d['reward'].long(), value=d['value'])
for floor in (False, True, False, True):
    ops.SPLIT_FLOOR = floor
    r = StepRunner(trainer, replay, B, use_graph=True)
    r.begin(torch.arange(B, device=dev), 0)
    for _ in range(6):
        r.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        r.step()
    e1.record()
    torch.cuda.synchronize()
    print(f'split rounding {"floor" if floor else "ceil"}: {e0.elapsed_time(e1) / 20:.3f} ms/step', flush=True)
