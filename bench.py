#!/usr/bin/env python
"""Benchmark of the SimPLe world-model hot path on B200 (contract: see the task description / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's own CPU path (baseline)

One "step" = one world-model optimiser step on one synthetic batch: NextFramePredictor forward (train branch), the
four/five losses, backward, clip_grad_norm_, Adafactor, and the scheduled-sampling frame feedback -- BASELINE.json
configs[1] ("World-model training ... bf16 on 1xB200, batch 64, with Adafactor").  `value` = global frames/s with the
replay resident in HBM; `e2e` = the same step fed from pinned host memory with the loss read back every step.
Also reported: simulated env-steps/s (configs[2]: 16 agents x 50 steps with a PPO policy), the tensor-core roofline of
the tcgen05 GEMM kernels, the HBM roofline of the fused softmax-CE kernel, and the reference's CPU number.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CE_DRAM_BYTES_B64 = 1599.7e6  # dram__bytes_read+write of one pixel_ce_kernel launch at B=64 (ncu, profiles/)
FUSED_CE_DRAM_BYTES_B64 = 872.2e6  # tap_gemm2_kernel<256,4,1> (logits GEMM + CE epilogue), same source, round 2
ALG_GFLOP_PER_FRAME_TRAIN = 36.80  # conv/deconv fwd+bwd per sample (algorithmic)


from simple_b200.synthetic import synth_replay  # noqa: E402  (synthetic "Atari-like" replay)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1400.0, 1590.0, 'fallback'


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the step (baseline/_ref copy, else the oracle port)
# ----------------------------------------------------------------------------------------------------------------
def reference_cpu_step_time(batch, steps, warmup, n_action=3):
    """Times `steps` world-model optimiser steps of the reference on the host CPU (fp32, all cores)."""
    import numpy as np
    from oracle import wm_oracle as O
    from oracle.ref_loader import load_reference
    torch.set_num_threads(os.cpu_count())
    cfg = O.default_config(device='cpu', batch_size=batch)
    ref = load_reference()
    data = synth_replay(batch + warmup + steps + 2, n_action, seed=0)
    torch.manual_seed(0)
    np.random.seed(0)
    times = []
    if ref is not None:
        kind = 'reference'
        model = ref.nfp.NextFramePredictor(cfg, n_action)
        opt = ref.adafactor.Adafactor(model.parameters())
        model.train()
        model.init_internal_states(batch)
        frames = data['obs'][:batch].float() / 255
        for j in range(warmup + steps):
            idx = torch.arange(batch) + j
            new_states = data['new_obs'][idx].long()
            t0 = time.perf_counter()
            logits, rew, val = model(frames, data['action'][idx].float(), new_states.float() / 255, 0)
            lstm = model.stochastic_model.get_lstm_loss()
            loss, *_ = O.wm_losses(cfg, logits, rew, val, lstm, new_states, data['reward'][idx].long(), data['value'][idx])
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.clip_grad_norm)
            opt.step()
            frames = torch.cat((frames[:, 3:], torch.argmax(logits.detach(), dim=1).float() / 255), dim=1)
            float(loss)
            if j >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = 'port'
        torch.manual_seed(0)
        from simple_b200.next_frame_predictor import NextFramePredictor
        p = {k: v.clone().requires_grad_(True) for k, v in NextFramePredictor(cfg, n_action).state_dict().items()}
        st = O.init_state(cfg, batch)
        opt_state = {k: {} for k in p}
        frames = data['obs'][:batch].float() / 255
        for j in range(warmup + steps):
            idx = torch.arange(batch) + j
            new_states = data['new_obs'][idx].long()
            t0 = time.perf_counter()
            logits, rew, val, lstm = O.forward(p, cfg, st, frames, data['action'][idx].float(), new_states.float() / 255,
                                               0, O.GlobalRand(), training=True)
            loss, *_ = O.wm_losses(cfg, logits, rew, val, lstm, new_states, data['reward'][idx].long(), data['value'][idx])
            names = list(p)
            gl = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
            present = [(k, g) for k, g in zip(names, gl) if g is not None]
            clipped, _ = O.clip_grad_norm([g for _, g in present], cfg.clip_grad_norm)
            with torch.no_grad():
                for (k, _), g in zip(present, clipped):
                    p[k].copy_(O.adafactor_step(p[k].detach(), g, opt_state[k]))
            frames = torch.cat((frames[:, 3:], torch.argmax(logits.detach(), dim=1).float() / 255), dim=1)
            if j >= warmup:
                times.append(time.perf_counter() - t0)
    return kind, sum(times) / len(times)


def reference_gpu_step_time(batch, n_action, dev, steps=4, warmup=3):
    """The UNMODIFIED reference (baseline/_ref) on the SAME GPU at the same batch: its own PyTorch-eager path (cuDNN /
    cuBLAS / ATen kernels, fp32 with PyTorch's default TF32 convolutions) for one full optimiser step.  Reported next to
    the CPU baseline so that the kernel-vs-library comparison is visible too."""
    import numpy as np
    from oracle import wm_oracle as O
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is None:
        return {'unavailable': 'reference not reachable'}
    cfg = O.default_config(device=str(dev), batch_size=batch)
    data = {k: v.to(dev) for k, v in synth_replay(batch + warmup + steps + 2, n_action, seed=0).items()}
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref.nfp.NextFramePredictor(cfg, n_action).to(dev)
    opt = ref.adafactor.Adafactor(model.parameters())
    model.train()
    model.init_internal_states(batch)
    frames = data['obs'][:batch].float() / 255
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for j in range(warmup + steps):
        if j == warmup:
            torch.cuda.synchronize()
            ev[0].record()
        idx = torch.arange(batch, device=dev) + j
        new_states = data['new_obs'][idx].long()
        logits, rew, val = model(frames, data['action'][idx].float(), new_states.float() / 255, 0)
        lstm = model.stochastic_model.get_lstm_loss()
        loss, *_ = O.wm_losses(cfg, logits, rew, val, lstm, new_states, data['reward'][idx].long(), data['value'][idx])
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.clip_grad_norm)
        opt.step()
        frames = torch.cat((frames[:, 3:], torch.argmax(logits.detach(), dim=1).float() / 255), dim=1)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    del model, opt
    torch.cuda.empty_cache()
    return {'value': batch / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms, 'batch': batch,
            'what': 'unmodified reference, device=cuda, PyTorch eager (cuDNN/cuBLAS, fp32 + default TF32 convs), '
                    f'{steps} optimiser steps after {warmup} warm-up steps'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 16  # bounded sample of the B=64 workload: one reference step on a quarter batch (a CPU step at 64 takes ~15 s)
    steps, warmup = max(1, min(args.steps, 4)), max(2, min(args.warmup, 3))
    kind, sec = reference_cpu_step_time(batch, steps, warmup)
    v = batch / sec
    line = {
        'impl': 'reference', 'metric': 'world_model_train_frames_per_s', 'value': v, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'world-model training step (fwd+loss+bwd+clip+Adafactor), Atari-100k shape 4x105x80x3, '
                               'batch 64 on the GPU arm; reference CPU arm times a bounded sample: batch 16 per step'},
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': kind,
                         'sample': f'{steps} optimiser steps at batch {batch} after {warmup} warm-up steps, fp32, '
                                   f'torch CPU backend, {os.cpu_count()} threads'},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# own arm
# ----------------------------------------------------------------------------------------------------------------
def run_own(args):
    # stdout carries the single JSON line; NCCL's log (whatever NCCL_DEBUG level the caller chose) goes to stderr
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    import torch.distributed as dist
    from simple_b200 import _lib, profiling
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.parallel import FlatGradSync, init_distributed
    from simple_b200.policy import CnnPolicy
    from simple_b200.simulated_env import Discrete, make_simulated_env
    from simple_b200.trainer import Trainer
    from types import SimpleNamespace

    torch.backends.cudnn.benchmark = True  # (the PyTorch stand-in policy's convolutions: let cuDNN pick its kernels)
    rank, world, local = init_distributed()
    dev = torch.device('cuda', local)

    def dbg(msg):
        if os.environ.get('SB_BENCH_DEBUG'):
            print(f'[bench rank {rank}] {msg}', file=sys.stderr, flush=True)
    torch.cuda.set_device(dev)
    _lib.check(_lib.lib().sb_init(local), 'sb_init')
    B, n_action, K, W_ = args.batch, args.n_action, args.steps, max(3, args.warmup)
    cfg = SimpleNamespace(
        agents=args.agents, batch_size=B, bottleneck_bits=128, bottleneck_noise=0.1, clip_grad_norm=1.0, compress_steps=5,
        device=str(dev), done_on_last_rollout_step=True, dropout=0.15, filter_double_steps=3, frame_shape=(3, 105, 80),
        hidden_layers=2, hidden_size=96, input_noise=0.05, latent_rnn_max_sampling=0.5, latent_state_size=128,
        latent_use_max_probability=0.8, recurrent_state_size=64, residual_dropout=0.5, rollout_length=50,
        save_models=False, scheduled_sampling_decay_steps=22250, stacking=4, stack_internal_states=True,
        target_loss_clipping=0.03, use_stochastic_model=True, use_wandb=False)
    torch.manual_seed(0)  # identical replicas
    model = NextFramePredictor(cfg, n_action).to(dev)
    trainer = Trainer(model, cfg)
    if world > 1:
        trainer.grad_sync = FlatGradSync(model)
    n_rec = B + max(K + W_ + 80, 2 * 50 + 30)
    host = synth_replay(n_rec, n_action, seed=1 + rank)
    data = {k: v.to(dev) for k, v in host.items()}
    pinned = {k: v.pin_memory() for k, v in host.items()}
    from simple_b200.trainer import StepRunner
    starts = torch.arange(B, device=dev)
    replay = dict(obs=data['obs'], new_obs=data['new_obs'], action=data['action'].float(), reward=data['reward'].long(),
                  value=data['value'])
    # the NCCL all-reduce of the data-parallel step is captured into the step graph as well (SB_DP_GRAPH=0: eager)
    use_graph = not args.no_graph and (world == 1 or os.environ.get('SB_DP_GRAPH', '1') != '0')
    runner = StepRunner(trainer, replay, B, use_graph=use_graph, resident=True)
    runner_host = StepRunner(trainer, replay, B, use_graph=use_graph, resident=False)
    pinned['action_f'] = host['action'].float().pin_memory()
    pinned['reward_l'] = host['reward'].long().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(r, n, j0, from_host):
        for j in range(n):
            if from_host:  # this step's batch: a contiguous slice of the pinned host replay, 4 async H2D copies
                k = j0 + j
                r.load_batch(pinned['new_obs'][k:k + B], pinned['action_f'][k:k + B], pinned['reward_l'][k:k + B],
                             pinned['value'][k:k + B])
            r.step()
            if from_host:
                float(r.last['loss'])  # device->host read of the step's result

    def timed(nsteps, from_host):
        r = runner_host if from_host else runner
        r.begin(starts, 0)
        if from_host:
            r.frames.copy_(trainer.preprocess_state(pinned['obs'][:B].to(dev, non_blocking=True), per_sample=False))
        run_steps(r, W_ + 3, 0, from_host)  # warm-up (first step of the window, graph warm-up + capture)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # launches = eager launches (host-side counter) + graph replays x kernels recorded in the captured step
        l0 = _lib.lib().sb_launch_count() + r.replays * r.graph_kernels
        e0.record()
        run_steps(r, nsteps, W_ + 3, from_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.lib().sb_launch_count() + r.replays * r.graph_kernels - l0, float(r.last['loss'])

    dbg('setup done')
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, last_loss = timed(K, False)
    clocks = sampler.finish() if sampler else None
    dbg('resident timed region done')
    ms_e2e, _, _ = timed(K, True)
    # ---- sustained: >= `--sustain-seconds` of FULL 50-step windows exactly as Trainer.train runs them (window start with
    # begin() = index sampling + input noise + state reset, the eager first step without the gate conv, then replays),
    # long enough for the clocks to settle under the power cap; the clock record is taken over this region
    sustained = None
    if args.sustain_seconds > 0:
        L_win = cfg.rollout_length
        smp = ClockSampler(local) if rank == 0 else None
        barrier()
        if smp:
            smp.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = time.perf_counter()
        e0.record()
        n_win = 0
        while True:
            runner.begin(starts, 0)
            for _ in range(L_win):
                runner.step()
            n_win += 1
            if n_win % 4 == 0:  # all ranks take the same decision (max over ranks of the elapsed host time)
                el = torch.tensor([time.perf_counter() - t_host], device=dev)
                if world > 1:
                    dist.all_reduce(el, op=dist.ReduceOp.MAX)
                if float(el) >= args.sustain_seconds:
                    break
        e1.record()
        barrier()
        ms_s = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t)
        sustained = {'value': B * world * n_win * L_win / (ms_s * 1e-3), 'unit': 'frames/s', 'seconds': ms_s * 1e-3,
                     'windows': n_win, 'steps': n_win * L_win, 'ms_per_step': ms_s / (n_win * L_win),
                     'includes': 'window begin(), eager first step of every window, graph replays',
                     'clocks': smp.finish() if smp else None}
    frames_per_s = B * world * K / (ms * 1e-3)
    e2e_fps = B * world * K / (ms_e2e * 1e-3)
    h2d = B * (3 * 105 * 80 + n_action + 1 + 4)
    line = {
        'metric': 'world_model_train_frames_per_s', 'value': frames_per_s, 'unit': 'frames/s', 'n_gpus': world,
        'steps': K, 'warmup': W_, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': 'world-model training step (NextFramePredictor train fwd + losses + bwd + clip_grad_norm '
                               '+ Adafactor + frame feedback), Atari-100k shape 4x105x80x3, n_action=%d' % n_action,
                   'per_gpu_batch': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush'},
        'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / K},
        'gpu_launches': int(launches), 'last_loss': last_loss, 'clocks': clocks, 'cuda_graph': bool(use_graph),
        'sustained': sustained,
    }

    dbg('timed regions done')
    if world > 1 and os.environ.get('SB_BENCH_DEBUG'):
        # replicas must stay identical: same reduced gradient, same deterministic update on every rank
        cs = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        both = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(both, cs)
        dbg('parameter checksums per rank: ' + ', '.join(f'{float(c):.10e}' for c in both))
    if rank == 0:
        hbm, tf_sus, tf_burst, how = peaks()
        # ---- roofline pass: per-launch CUDA-event timing of the hot kernels over 3 extra steps.  Rank 0 only, so the
        # gradient all-reduce is switched off for it (a collective entered by one rank alone would never return).
        sync_saved, trainer.grad_sync = trainer.grad_sync, None
        scale_saved, trainer.optimizer.grad_scale = trainer.optimizer.grad_scale, 1.0  # local gradients: no DP sum
        eager = StepRunner(trainer, replay, B, use_graph=False, resident=True)
        eager.begin(starts, 0)
        for j in range(2):
            eager.step()
        torch.cuda.synchronize()
        profiling.RECORDER = rec = profiling.Recorder()
        for j in range(3):
            eager.step()
        summ = rec.summary()
        profiling.RECORDER = None
        trainer.grad_sync = sync_saved
        trainer.optimizer.grad_scale = scale_saved
        gem = [summ[k] for k in ('tap_gemm_kernel', 'tap_wgrad_kernel') if k in summ]
        if gem:
            fl, tm, nl = sum(g['work'] for g in gem), sum(g['ms'] for g in gem), sum(g['launches'] for g in gem)
            ach = fl / (tm * 1e-3) / 1e12
            # the launches are timed one by one in a short eager pass at full clocks: the denominator is the BURST peak
            pure = [(t_, v_) for g in gem for t_, v_ in g['by_tag'].items() if '+ce' not in t_]
            fl_p, tm_p = sum(v_[1] for _, v_ in pure), sum(v_[0] for _, v_ in pure)
            line['roofline'] = {'bound': 'tensor', 'kernel': 'tap_gemm_kernel+tap_gemm2_kernel+tap_wgrad_kernel (tcgen05)',
                                'achieved': ach, 'peak': tf_burst, 'unit': 'TFLOP/s', 'frac': ach / tf_burst,
                                'frac_of_sustained_peak': ach / tf_sus, 'traffic': None,
                                'peak_source': f'{how} burst bf16 (per-launch CUDA events at full clocks)',
                                'launches_per_step': nl / 3, 'ms_per_step': tm / 3,
                                'alg_gflop_per_step': fl / 3 / 1e9,
                                'note': 'includes the logits GEMM whose epilogue carries the softmax-CE (its whole '
                                        'duration is charged to the GEMM flops); without that launch: '
                                        f'{fl_p / (tm_p * 1e-3) / 1e12 / tf_burst:.3f} of the burst peak'}
            line['roofline_by_layer'] = {
                t: {'ms': v[0] / 3, 'tflops': (v[1] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else None}
                for g in gem for t, v in sorted(g['by_tag'].items())}
        ce_key = 'logits_ce_kernel' if 'logits_ce_kernel' in summ else 'pixel_ce_kernel'
        if ce_key in summ:
            c = summ[ce_key]
            ach = c['work'] / (c['ms'] * 1e-3) / 1e9
            fused = ce_key == 'logits_ce_kernel'
            line['roofline_ce'] = {
                'bound': 'hbm',
                'kernel': ('tap_gemm2_kernel<256,4,1>: logits GEMM + softmax-CE epilogue (the logits never reach HBM)'
                           if fused else 'pixel_ce_kernel'),
                'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'traffic': (FUSED_CE_DRAM_BYTES_B64 if fused else CE_DRAM_BYTES_B64) if B == 64 else None,
                'traffic_source': ('ncu dram__bytes_read.sum + dram__bytes_write.sum of the launch, '
                                   'profiles/r2_ncu_launches_train_b64_v8.csv (B=64)') if fused
                else 'ncu --set full, profiles/r1_ncu_full_v13.summary.txt (B=64)',
                'peak_source': how, 'ms_per_launch': c['ms'] / c['launches'],
                'alg_bytes_per_launch': c['work'] / c['launches'],
                'alg_bytes_definition': ('read hf (bf16 x 96) + write dlogits (bf16 x 768) + targets + arg-max per pixel; '
                                         'the unfused pair (logits GEMM, pixel_ce_kernel) moved 3 x 768 bf16 per pixel')
                if fused else 'read logits + write dlogits (bf16 x 768 each) + targets + arg-max per pixel'}
    # ---- simulated rollout: agents x 50 steps with a PPO-style CNN policy (configs[2]); agents sharded over ranks
    A = args.agents
    cfg_env = SimpleNamespace(**{**vars(cfg), 'agents': A})
    env = make_simulated_env(cfg_env, model, Discrete(n_action))
    policy = CnnPolicy((12, 105, 80), n_action).to(dev)
    env.restart_all(data['obs'][:A])

    env.use_graph = not args.no_graph
    env.attach_policy(policy)  # policy.act is part of the captured step (sampled on the device from the last obs)

    def rollout():
        env.reset()
        for _ in range(cfg.rollout_length):
            obs, rew, done, _ = env.step(None)
        return rew

    dbg('roofline pass done')
    rollout()
    barrier()
    dbg('rollout warm-up done')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.rollouts):
        rollout()
    e1.record()
    barrier()
    rms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([rms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rms = float(t)
    ms_env = rms / (cfg.rollout_length * args.rollouts)
    line['rollout'] = {'metric': 'simulated_env_steps_per_s', 'value': A * world * cfg.rollout_length * args.rollouts / (rms * 1e-3),
                       'unit': 'env-steps/s', 'agents_per_gpu': A, 'n_action': n_action, 'steps': cfg.rollout_length,
                       'ms_per_env_step_batch': ms_env,
                       'includes': 'reset + eager first step of every rollout, PPO-style CNN policy act inside the step'}
    if rank == 0:
        hbm, tf_sus, tf_burst, how = peaks()
        # eval forward of the world model: 11.68 GFLOP per agent-step on the tensor cores; at 16 agents the
        # step is latency-bound (every layer is a sub-wave launch), which this fraction shows
        gf = 11.68 * A
        line['rollout']['roofline'] = {'bound': 'tensor', 'achieved': gf / ms_env, 'peak': tf_burst,
                                       'unit': 'TFLOP/s', 'frac': gf / ms_env / tf_burst,
                                       'alg_gflop_per_env_step_batch': gf, 'peak_source': f'{how} burst bf16'}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:  # the unmodified reference stack: make_simulated_env (16 subprocesses) + PPO policy on the host CPU
            r = subprocess.run([sys.executable, '-m', 'oracle.ref_rollout', '--agents', str(A), '--steps', '8',
                                '--warmup', '2', '--n-action', str(n_action)], cwd=REPO, capture_output=True, text=True,
                               timeout=600)
            jl = [l for l in r.stdout.splitlines() if l.startswith('{')]
            rb = json.loads(jl[-1]) if jl else {'unavailable': (r.stderr or 'no output')[-200:]}
        except Exception as e:  # noqa: BLE001
            rb = {'unavailable': str(e)[:200]}
        if 'env_steps_per_s' in rb:
            line['rollout']['cpu_baseline'] = {
                'value': rb['env_steps_per_s'], 'unit': 'env-steps/s', 'cores': rb['cores'], 'kind': 'reference',
                'sample': f"{rb['steps']} env steps x {rb['agents']} agents after 2 warm-up steps: reference "
                          'make_simulated_env (SubprocVecEnv) + PPO policy, device=cpu, fp32'}
        else:
            line['rollout']['cpu_baseline'] = rb
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, sec = reference_cpu_step_time(16, 2, 2)
        line['cpu_baseline'] = {'value': 16 / sec, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': kind,
                                'sample': '2 optimiser steps at batch 16 after 2 warm-up steps (gate conv active), fp32, '
                                          f'torch CPU backend, {os.cpu_count()} threads'}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line['reference_on_gpu'] = reference_gpu_step_time(B, n_action, dev)
        except Exception as e:  # noqa: BLE001
            line['reference_on_gpu'] = {'unavailable': str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured CUDA graphs hold NCCL work: tearing the process group down underneath them can block forever, so the
        # ranks meet at a barrier, drain the device and leave without running the communicator's destructor.
        dbg('leaving')
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='own', choices=['own', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch (weak scaling)')
    ap.add_argument('--agents', type=int, default=16, help='simulated agents per GPU')
    ap.add_argument('--rollouts', type=int, default=2)
    ap.add_argument('--n-action', type=int, default=3, help='3 = Freeway (configs[1-3]); 9 = MsPacman (configs[4])')
    ap.add_argument('--sustain-seconds', type=float, default=5.0,
                    help='length of the sustained full-window measurement (0 = skip)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run every step eagerly (no CUDA graph replay)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
