"""Which Python lines of the training step launch the small ATen kernels?  One eager step under torch.profiler with
stacks; device time and launch count per (simple_b200 source line, aten op).  python tools/small_ops.py [batch]"""
import os
import sys
import collections
import torch
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_replay  # noqa: E402
from simple_b200.next_frame_predictor import NextFramePredictor  # noqa: E402
from simple_b200.trainer import StepRunner, Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
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
d = {k: v.to(dev) for k, v in synth_replay(B + 20, 3, 1).items()}
replay = dict(obs=d['obs'], new_obs=d['new_obs'], action=d['action'].float(), reward=d['reward'].long(), value=d['value'])
r = StepRunner(trainer, replay, B, use_graph=False)
r.begin(torch.arange(B, device=dev), 0)
for _ in range(3):
    r.step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            with_stack=True, record_shapes=True,
                            experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    r.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if not e.name.startswith('aten::') or e.device_type != torch.autograd.DeviceType.CPU:
        continue
    ks = [k for k in e.kernels]
    if not ks:
        continue
    t = sum(k.duration for k in ks)
    site = '?'
    for fr in (e.stack or []):
        if 'simple_b200/' in fr and 'profiling.py' not in fr:
            site = fr.split('simple_b200/')[-1]
            break
    if site == '?':
        site = 'autograd/other ' + str(e.input_shapes)[:60]
    a = agg[(site, e.name)]
    a[0] += len(ks)
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f'aten kernels: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms')
by_site = collections.defaultdict(lambda: [0, 0.0])
for (site, op), (n, t) in agg.items():
    by_site[site][0] += n
    by_site[site][1] += t
print('--- by source line')
for site, (n, t) in sorted(by_site.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f'{t:8.1f} us {n:4d}x  {site}')
print('--- by (line, op)')
for (site, op), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f'{t:8.1f} us {n:4d}x  {op:28s} {site}')
