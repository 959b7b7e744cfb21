"""Kernels of the replayed simulated-env step (CUPTI): time per kernel and, with SB_DUMP=path, every kernel of the last
step in order.  Usage on the GPU box: python tools/rollout_gaps.py [agents] [n_action]"""
import os
import sys
import torch
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_replay  # noqa: E402
from simple_b200.next_frame_predictor import NextFramePredictor  # noqa: E402
from simple_b200.policy import CnnPolicy  # noqa: E402
from simple_b200.simulated_env import Discrete, make_simulated_env  # noqa: E402

A = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_action = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
cfg = SimpleNamespace(
    agents=A, batch_size=A, bottleneck_bits=128, bottleneck_noise=0.1, clip_grad_norm=1.0, compress_steps=5,
    device=str(dev), done_on_last_rollout_step=True, dropout=0.15, filter_double_steps=3, frame_shape=(3, 105, 80),
    hidden_layers=2, hidden_size=96, input_noise=0.05, latent_rnn_max_sampling=0.5, latent_state_size=128,
    latent_use_max_probability=0.8, recurrent_state_size=64, residual_dropout=0.5, rollout_length=50,
    save_models=False, scheduled_sampling_decay_steps=22250, stacking=4, stack_internal_states=True,
    target_loss_clipping=0.03, use_stochastic_model=True, use_wandb=False)
torch.manual_seed(0)
model = NextFramePredictor(cfg, n_action).to(dev)
data = {k: v.to(dev) for k, v in synth_replay(A + 20, n_action, 1).items()}
env = make_simulated_env(cfg, model, Discrete(n_action))
policy = CnnPolicy((12, 105, 80), n_action).to(dev)
env.restart_all(data['obs'][:A])
env.use_graph = True
env.attach_policy(policy)
env.reset()
for _ in range(10):
    env.step(None)
torch.cuda.synchronize()
N = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(N):
        env.step(None)
    e1.record()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
dur = lambda e: (e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total)
print(f'agents={A}: wall {e0.elapsed_time(e1) / N:.3f} ms/step, GPU kernel time {sum(dur(e) for e in evs) / 1e3 / N:.3f} ms/step, '
      f'{len(evs) / N:.0f} kernels/step')
agg = {}
for e in evs:
    a = agg.setdefault(e.name.split('(')[0][:90], [0, 0.0])
    a[0] += 1 / N
    a[1] += dur(e) / 1e3 / N
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('SB_TOP', '45'))]:
    print(f'{t:8.4f} ms {n:6.1f}x  {k}')
own = sum(t for k, (n, t) in agg.items() if 'sb::' in k)
print(f'--- sb:: kernels {own:.3f} ms, others {sum(t for k, (n, t) in agg.items() if "sb::" not in k):.3f} ms '
      f'({sum(n for k, (n, t) in agg.items() if "sb::" not in k):.0f} launches)')
if os.environ.get('SB_DUMP'):
    tl = sorted(((e.time_range.start, e.time_range.end, e.name.split('(')[0][:110]) for e in evs), key=lambda x: x[0])
    last = tl[-(len(tl) // N):]
    with open(os.environ['SB_DUMP'], 'w') as f:
        for a, b, n in last:
            f.write(f'{a - last[0][0]:9.1f} {b - a:7.1f}  {n}\n')
