"""GPU busy time vs wall time of the CUDA-graph-replayed training step (torch.profiler / CUPTI), plus the kernels that
take the most time inside the graph.  Usage on the GPU box: python tools/graph_gaps.py [batch]"""
import os
import sys
import torch
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_replay  # noqa: E402
from simple_b200.next_frame_predictor import NextFramePredictor  # noqa: E402
from simple_b200.trainer import StepRunner, Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
world = int(os.environ.get('WORLD_SIZE', '1'))
rank, local = 0, 0
if world > 1:  # under torchrun: data-parallel step with the overlapped gradient all-reduce; rank 0 reports
    from simple_b200.parallel import FlatGradSync, init_distributed
    rank, world, local = init_distributed()
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
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
if world > 1:
    trainer.grad_sync = FlatGradSync(model)
d = {k: v.to(dev) for k, v in synth_replay(B + 40, 3, 1 + rank).items()}
replay = dict(obs=d['obs'], new_obs=d['new_obs'], action=d['action'].float(), reward=d['reward'].long(), value=d['value'])
r = StepRunner(trainer, replay, B, use_graph=True)
r.begin(torch.arange(B, device=dev), 0)
for _ in range(8):
    r.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get('SB_NOPROF'):  # plain timing (no CUPTI): median of 5 x 30 replayed steps, max over ranks
    ts = []
    for _ in range(5):
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        r.begin(torch.arange(B, device=dev), 0)  # the synthetic replay holds 40 steps per window
        for _ in range(3):
            r.step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            r.step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 30)
    t = torch.tensor(sorted(ts)[2], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'B={B}: wall {float(t):.3f} ms/step (median of 5 x 30 steps, no profiler; all: {[round(x, 3) for x in ts]})')
    os._exit(0)
N = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(N):
        r.step()
    e1.record()
    torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / N
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()
    if rank != 0:
        os._exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total for e in evs) / 1e3 / N
print(f'B={B}: wall {wall:.3f} ms/step, GPU kernel time {busy:.3f} ms/step, {len(evs) / N:.0f} kernels/step')
agg = {}
for e in evs:
    t = (e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total) / 1e3 / N
    k = e.name.split('(')[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1 / N
    a[1] += t
top = int(os.environ.get('SB_TOP', '40'))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{t:8.3f} ms {n:6.1f}x  {k}')
own = [(n, t) for k, (n, t) in agg.items() if 'sb::' in k]
other = [(n, t) for k, (n, t) in agg.items() if 'sb::' not in k]
print(f'--- kernels of this library (sb::): {sum(n for n, _ in own):.0f} launches, {sum(t for _, t in own):.3f} ms; '
      f'library / ATen kernels: {sum(n for n, _ in other):.0f} launches, {sum(t for _, t in other):.3f} ms')
# timeline of the last profiled step: idle gaps of the GPU (> 15 us) and where the NCCL kernels sit
tl = sorted(((e.time_range.start, e.time_range.end, e.name.split('(')[0][:110]) for e in evs), key=lambda x: x[0])
per = len(tl) // N
last = tl[-per:]
t0 = last[0][0]
print(f'--- last step timeline: {len(last)} kernels, span {(max(x[1] for x in last) - t0) / 1e3:.3f} ms')
end = last[0][1]
prev = last[0][2]
for a, b, n in last[1:]:
    if a - end > 15:
        print(f'  gap {a - end:7.1f} us at +{(end - t0) / 1e3:6.3f} ms  after {prev}  before {n}')
    if b > end:
        end, prev = b, n
for a, b, n in last:
    if 'nccl' in n:
        print(f'  nccl +{(a - t0) / 1e3:6.3f} .. +{(b - t0) / 1e3:6.3f} ms  ({(b - a):7.1f} us)')
for key in ('pixel_ce_kernel', 'adafactor_tick', 'grad_sumsq', 'frame_feedback'):
    for a, b, n in last:
        if key in n:
            print(f'  mark {key} +{(a - t0) / 1e3:6.3f} ms')
            break
if os.environ.get('SB_DUMP'):  # every kernel of the last replayed step, in order: start (us), duration (us), name
    with open(os.environ['SB_DUMP'], 'w') as f:
        for a, b, n in last:
            f.write(f'{a - t0:9.1f} {b - a:7.1f}  {n}\n')
if world > 1:
    sys.stdout.flush()
    os._exit(0)
