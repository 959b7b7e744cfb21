"""World-model trainer (drop-in for simple/trainer.py of the reference): Trainer(model, config).train(epoch, env, steps).

Same loop semantics (windows of `rollout_length` consecutive transitions, input noise with median replacement,
scheduled sampling in epoch 0, four/five losses, clip_grad_norm_, Adafactor, one optimiser step per time step) but
device-resident: the replay list is mirrored once into contiguous device tensors, windows are gathered on the device,
losses are accumulated on the device and read back once per window, and clip+Adafactor are fused CUDA kernels.
"""
import os

import torch

from . import ops
from .adafactor import Adafactor

NVTX = os.environ.get('SB_NVTX', '0') == '1'  # NVTX ranges around forward+loss / backward / grad sync / optimiser


_ONES = {}


def _one(device, dtype):
    """Persistent constant 1 (gradient seeds: no fill kernel per step, valid under CUDA-graph capture)."""
    key = (torch.device(device), dtype)
    if key not in _ONES:
        _ONES[key] = torch.ones((1,) if dtype == torch.float64 else (), device=device, dtype=dtype)
    return _ONES[key]


class TotalLoss(torch.autograd.Function):
    """loss = loss_reconstruct + loss_value + loss_reward (+ loss_lstm) (simple/trainer.py:134-144) from the kernels' raw
    results in ONE launch (sb_loss_total), optionally added to a running [5] log accumulator.  Every summand enters with
    weight 1: the backward pass hands the incoming gradient through (the producers computed their gradients for weight 1
    in their forward launches)."""

    @staticmethod
    def forward(ctx, rec_sum, n, clip, l_value, l_reward, l_lstm, running):
        from . import _lib
        out = torch.empty(5, device=rec_sum.device, dtype=torch.float32)
        _lib.check(_lib.lib().sb_loss_total(rec_sum.data_ptr(), float(n), float(clip), l_value.data_ptr(),
                                            l_reward.data_ptr(), l_lstm.data_ptr() if l_lstm is not None else None,
                                            out.data_ptr(), running.data_ptr() if running is not None else None,
                                            _lib.stream_ptr()), 'sb_loss_total')
        ctx.has_lstm = l_lstm is not None
        ctx.dev = rec_sum.device
        parts = tuple(out[i] for i in range(5))
        ctx.mark_non_differentiable(*parts[1:])
        return parts

    @staticmethod
    def backward(ctx, g, *_):
        return _one(ctx.dev, torch.float64), None, None, g, g, (g if ctx.has_lstm else None), None


def wm_loss(model, config, frames, actions, new_states_u8, rewards, values, epsilon, keep_logits=False, loss_acc=None):
    """Forward + the losses of simple/trainer.py:134-144.  Returns a dict of device scalars (no host sync).

    frames f32 [B,12,H,W] in [0,1]; actions f32 [B,n] one-hot; new_states_u8 u8 [B,3,H,W]; rewards long [B];
    values f32 [B].  The pixel cross-entropy, its gradient and the argmax decode are ONE fused kernel pass over the
    bf16 logits (ops.PixelCE); `argmax` is the u8 frame the reference obtains with torch.argmax (trainer.py:130).
    """
    target = new_states_u8 / 255  # (true division of the u8 frame: float32, same values as .float() / 255)
    new_states_u8 = new_states_u8.contiguous()
    clip = float(config.target_loss_clipping)
    out = {}
    if keep_logits:  # tests: materialised logits + the stand-alone fused loss kernel (which overwrites them with their gradient)
        logits, reward_pred, value_pred = model.forward_internal(frames, actions, target, epsilon, rewards=rewards,
                                                                 values=values)
        B, H, W, _ = logits.shape
        out['logits_ref'] = logits.detach().reshape(B, H, W, 3, 256).permute(0, 4, 3, 1, 2).float().clone()
        loss_reconstruct, argmax = ops.PixelCE.apply(logits, new_states_u8, clip, True, True,
                                                     getattr(model, 'logits_dbias', None))
    else:  # the logits GEMM carries the loss in its epilogue: the logits are never written to HBM
        _, reward_pred, value_pred = model.forward_internal(frames, actions, target, epsilon, rewards=rewards,
                                                            values=values, logits_mode='ce_raw', ce_target=new_states_u8,
                                                            ce_clip=clip)
        aux = model.last_aux
        l_lstm = model.stochastic_model.get_lstm_loss() if config.use_stochastic_model else None
        loss, l_rec, l_val, l_rew, l_lstm_out = TotalLoss.apply(aux['rec_sum'], aux['rec_n'], clip, aux['loss_value'],
                                                                aux['loss_reward'], l_lstm, loss_acc)
        if config.use_stochastic_model:
            out['loss_lstm'] = l_lstm_out
        out.update(loss=loss, loss_reconstruct=l_rec, loss_value=l_val, loss_reward=l_rew, argmax=aux['argmax'],
                   reward_pred=reward_pred, value_pred=value_pred)
        return out
    # value MSE and reward CE (trainer.py:138-139) come out of the fused heads kernel together with their gradients
    loss_value, loss_reward = model.last_aux['loss_value'], model.last_aux['loss_reward']
    loss = loss_reconstruct + loss_value + loss_reward
    if config.use_stochastic_model:
        loss_lstm = model.stochastic_model.get_lstm_loss()
        loss = loss + loss_lstm
        out['loss_lstm'] = loss_lstm
    out.update(loss=loss, loss_reconstruct=loss_reconstruct, loss_value=loss_value, loss_reward=loss_reward,
               argmax=argmax, reward_pred=reward_pred, value_pred=value_pred)
    return out


class DeviceReplay:
    """Contiguous device mirror of the reference's replay list (atari_utils/atari_utils/envs.py:149-156):
    records [obs u8[12,H,W], action u8[n] one-hot, reward u8, new_obs u8[3,H,W], done u8, value f32|None]."""

    def __init__(self, buffer, device, rollout_len):
        n = len(buffer)
        self.n = n
        self.obs = torch.stack([r[0] for r in buffer]).to(device)
        self.action = torch.stack([r[1] for r in buffer]).to(device).float()
        self.reward = torch.stack([r[2].reshape(()) for r in buffer]).to(device).long()
        self.new_obs = torch.stack([r[3] for r in buffer]).to(device)
        done = torch.tensor([bool(r[4]) for r in buffer])
        has_value = torch.tensor([r[5] is not None for r in buffer])
        self.value = torch.stack([r[5].reshape(()).float().cpu() if r[5] is not None else torch.zeros(())
                                  for r in buffer]).to(device)
        # a start index is valid when none of its rollout_len successors is terminal or lacks a value (trainer.py:44-53)
        bad = (done | ~has_value).long()
        csum = torch.cat((torch.zeros(1, dtype=torch.long), torch.cumsum(bad, 0)))
        starts = torch.arange(0, n - rollout_len)
        ok = (csum[starts + rollout_len] - csum[starts]) == 0
        self.valid_starts = starts[ok].to(device)

    def sample_starts(self, batch):
        if self.valid_starts.numel() == 0:
            raise BufferError('Can\'t train the world model, the buffer does not contain one full rollout.')
        idx = torch.randint(self.valid_starts.numel(), (batch,), device=self.valid_starts.device)
        return self.valid_starts[idx]


class Trainer:

    def __init__(self, model, config):
        self.model = model
        self.config = config
        self.logger = None
        if self.config.use_wandb:
            # The reference logs through its own atari_utils.logger (out of scope here): any object with `.log(dict)` can
            # be assigned to `trainer.logger`; nothing of the reference is imported by this package.
            raise ValueError('simple_b200.Trainer: use_wandb is not supported; assign an object with .log(dict) to '
                             'trainer.logger instead')
        self.optimizer = Adafactor(self.model.parameters())
        self.model_step = 1
        self.reward_step = 1
        self._replay = None
        self._replay_key = None
        self.grad_sync = None  # set by simple_b200.parallel for data-parallel training
        # conv-weight gradients stay in the packed GEMM-operand layout between wgrad and Adafactor (no unpack / re-pack
        # passes); `p.grad` of those weights is then None after a step.  Set False to get reference-shaped p.grad.
        self.packed_grads = True

    # ---- simple/trainer.py:58-65
    def preprocess_state(self, state_u8, per_sample):
        state = state_u8.float() / 255
        p = self.config.input_noise
        if p <= 0:
            return state
        keep = (torch.rand(state.shape, device=state.device) >= p).to(state.dtype)
        if per_sample:  # one frame [3,H,W] at a time in the reference => one median per sample
            med = state.reshape(state.shape[0], -1).median(dim=1).values.reshape(-1, 1, 1, 1)
        else:
            med = self.global_median(state_u8)
        return state * keep + med * (1 - keep)

    def global_median(self, state_u8):
        """torch.median(state / 255) over the WHOLE batch (simple/trainer.py:64) as a device scalar.  Under data
        parallelism the batch is sharded: the 256-bin histograms of the shards are all-reduced, and the lower median
        (the element of rank (N-1)//2, what torch.median returns) is read off the prefix sums -- exactly the median of
        the global batch (SURVEY 8e caveat 1), not a per-shard one."""
        import torch.distributed as dist
        if not (self.grad_sync is not None and dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1):
            return torch.median(state_u8.float() / 255)
        hist = torch.bincount(state_u8.reshape(-1).long(), minlength=256)
        dist.all_reduce(hist)
        k = (hist.sum() - 1) // 2
        return (torch.cumsum(hist, 0) <= k).sum().float() / 255

    def epsilon_schedule(self, epoch, i):
        """Probability of feeding the ground-truth frame (trainer.py:78-87)."""
        if epoch != 0:
            return 0
        decay_steps = self.config.scheduled_sampling_decay_steps
        inv_base = torch.exp(torch.log(torch.tensor(0.01)) / (decay_steps // 4))
        epsilon = inv_base ** max(decay_steps // 4 - i, 0)
        progress = min(i / decay_steps, 1)
        progress = progress * (1 - 0.01) + 0.01
        epsilon *= progress
        return float(1 - epsilon)

    # ---- checkpoint / resume.  The reference saves the model weights only (trainer.py:175-176) and
    # restarts the optimiser from scratch; this also carries the Adafactor state, the step counters and the RNG streams.
    def save_checkpoint(self, path, epoch=0):
        opt_state = []
        for p in self.model.parameters():
            st = self.optimizer.state.get(p, {})
            opt_state.append({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in st.items()
                              if k in ('step', 'exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq')})
        seed = self.model._seed_dev
        torch.save({'model': {k: v.detach().cpu() for k, v in self.model.state_dict().items()}, 'optimizer': opt_state,
                    'grad_scale': self.optimizer.grad_scale,  # units of the second-moment states (1 / DP world size)
                    'model_step': self.model_step, 'reward_step': self.reward_step, 'epoch': int(epoch),
                    'dropout_seed': int(seed.item()) if seed is not None else None,
                    'torch_rng': torch.get_rng_state(),
                    'cuda_rng': torch.cuda.get_rng_state(self.model.logits.weight.device)}, path)

    def load_checkpoint(self, path):
        """Restores weights, optimiser state, counters and RNG streams; returns the stored epoch."""
        ck = torch.load(path, map_location='cpu')
        self.model.load_state_dict(ck['model'])  # (also invalidates the packed bf16 operands)
        dev = self.model.logits.weight.device
        self.optimizer.state.clear()
        self.optimizer._cache.clear()
        current_scale = self.optimizer.grad_scale
        self.optimizer._grad_scale = float(ck.get('grad_scale', 1.0))  # the stored states' units ...
        for p, saved in zip(self.model.parameters(), ck['optimizer']):
            if not saved:
                continue
            st = self.optimizer._init_state(p)
            st['step'] = int(saved['step'])
            st['step_dev'].fill_(float(saved['step']))
            for k in ('exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq'):
                if k in saved:
                    st[k].copy_(saved[k].to(dev))
        self.optimizer.grad_scale = current_scale  # ... converted to this run's (another world size)
        self.model_step, self.reward_step = ck['model_step'], ck['reward_step']
        if ck.get('dropout_seed') is not None:
            if self.model._seed_dev is not None and self.model._seed_dev.device == dev:
                self.model._seed_dev.fill_(ck['dropout_seed'])  # in place: captured graphs read this address
            else:
                self.model._seed_dev = torch.full((1,), ck['dropout_seed'], dtype=torch.int64, device=dev)
        torch.set_rng_state(ck['torch_rng'])
        torch.cuda.set_rng_state(ck['cuda_rng'], dev)
        self._runner = None  # captured graphs refer to the old optimiser records
        return ck['epoch']

    def train_step(self, frames, actions, new_states_u8, rewards, values, epsilon, after_backward=None, loss_acc=None):
        """One optimiser step (trainer.py:122-149).  Returns the loss dict of `wm_loss` (device scalars).
        `after_backward(out)`: optional gradient-independent tail work of the caller (frame feedback); it is issued
        between the backward pass and the optimiser so that it overlaps the last gradient all-reduce."""
        self.model.train()
        nvtx = torch.cuda.nvtx if NVTX else None  # ranges for Nsight timelines (host-side markers: graph-capture safe)
        if nvtx:
            nvtx.range_push('wm_step/forward+loss')
        ops.ARENA.begin(frames.device)  # one memset for all zero-initialised scratch of the step
        out = wm_loss(self.model, self.config, frames, actions, new_states_u8, rewards, values, epsilon,
                      loss_acc=loss_acc)
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push('wm_step/backward')
        self.optimizer.zero_grad(set_to_none=True)
        packed = self.model.packed_weight_map(defer_unpack=self.packed_grads) if self.packed_grads else None
        if self.grad_sync is not None and hasattr(self.grad_sync, 'prepare'):
            self.grad_sync.prepare(self.model, packed)  # gradient arena + bucketed all-reduce overlapped with backward
        out['loss'].backward(gradient=_one(out['loss'].device, torch.float32))

        def tail():
            self.model.commit_state()  # after backward: the gate's weight gradient still needed the previous operand
            if after_backward is not None:
                after_backward(out)
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push('wm_step/grad_sync+feedback')
        if self.grad_sync is not None:
            grads = self.grad_sync(self.model, packed, tail)
            # overlap mode hands back the SUM over the ranks (see Adafactor.grad_scale); set outside graph capture
            self.optimizer.grad_scale = getattr(self.grad_sync, 'last_scale', 1.0)
        else:
            grads = None
            tail()
        if nvtx:
            nvtx.range_pop()
            nvtx.range_push('wm_step/clip+adafactor')
        self.optimizer.step(max_grad_norm=self.config.clip_grad_norm, grads=grads, packed=packed)
        if nvtx:
            nvtx.range_pop()
        if packed:
            for pw in packed.values():
                pw.gpacked = None
        ops.ARENA.end()
        return out

    def train(self, epoch, env, steps=15000):
        if epoch == 0:
            steps *= 3
        cfg = self.config
        rollout_len = cfg.rollout_length
        B = cfg.batch_size
        dev = torch.device(cfg.device)
        if env.buffer[0][5] is None:
            raise BufferError('Can\'t train the world model, the buffer does not contain one full episode.')
        states, actions, rewards, new_states, dones, values = env.buffer[0]
        assert states.dtype == torch.uint8
        assert actions.dtype == torch.uint8
        assert rewards.dtype == torch.uint8
        assert new_states.dtype == torch.uint8
        assert values.dtype == torch.float32
        key = (id(env.buffer), len(env.buffer))
        if self._replay_key != key:
            self._replay = DeviceReplay(env.buffer, dev, rollout_len)
            self._replay_key = key
            self._runner = None
        rp = self._replay
        if getattr(self, '_runner', None) is None:
            self._runner = StepRunner(self, dict(obs=rp.obs, new_obs=rp.new_obs, action=rp.action, reward=rp.reward,
                                                 value=rp.value), B, use_graph=getattr(cfg, 'cuda_graph', True))
        runner = self._runner
        try:
            from tqdm import trange
            iterator = trange(0, steps, rollout_len, desc='Training world model', unit_scale=rollout_len)
        except ImportError:
            iterator = range(0, steps, rollout_len)
        metrics = {}
        for i in iterator:
            epsilon = self.epsilon_schedule(epoch, i)
            # (data parallel: no re-synchronisation of the replicas is needed -- every reduction downstream of the gradient
            # all-reduce is order-deterministic, so identical replicas stay bit-identical; tests/test_dp_nccl.py)
            runner.begin(rp.sample_starts(B), epsilon)
            for j in range(rollout_len):
                runner.step()
            host = (runner.losses / rollout_len).tolist()  # the only host sync of the window
            metrics = dict(zip(runner.names, host))
            if self.logger is not None:
                d = {'model_step': self.model_step, 'epsilon': epsilon}
                d.update(metrics)
                self.logger.log(d)
                self.model_step += rollout_len
            if hasattr(iterator, 'set_postfix'):
                iterator.set_postfix(metrics)
        torch.cuda.empty_cache()
        if cfg.save_models:
            torch.save(self.model.state_dict(), os.path.join('models', 'model.pt'))
        return metrics


class StepRunner:
    """Runs the optimiser steps of one window (simple/trainer.py:104-154) on static device buffers.

    The first step of a window differs (no gate convolution, gate.* without gradient) and runs eagerly; every other
    step is identical work on identical addresses, so it is captured ONCE into a CUDA graph (forward, fused loss,
    backward, clip + Adafactor, frame feedback, state commit, index advance) and replayed: one launch per step instead
    of ~1500, which removes the host from the critical path (Blackwell guide, guideline 9).
    """

    def __init__(self, trainer, replay, B, use_graph=True, resident=True):
        self.trainer, self.replay, self.B = trainer, replay, B
        cfg = trainer.config
        dev = torch.device(cfg.device)
        c, h, w = cfg.frame_shape
        self.c = c
        self.use_graph = use_graph
        self.resident = resident
        self.idx = torch.zeros(B, dtype=torch.long, device=dev)
        self.frames = torch.zeros(B, c * int(cfg.stacking), h, w, device=dev)
        self.eps = torch.zeros((), device=dev)
        n_losses = 5 if cfg.use_stochastic_model else 4
        self.names = ['loss', 'loss_reconstruct', 'loss_value', 'loss_reward', 'loss_lstm'][:n_losses]
        self.losses = torch.zeros(n_losses, device=dev)
        self.inp = dict(new_obs=torch.zeros(B, c, h, w, dtype=torch.uint8, device=dev),
                        action=torch.zeros(B, replay['action'].shape[1], device=dev),
                        reward=torch.zeros(B, dtype=torch.long, device=dev), value=torch.zeros(B, device=dev))
        self.graph = None
        self.graph_key = None
        self.model_key = None
        self.graph_kernels = 0  # kernels of this library recorded in the captured step
        self.replays = 0
        self.warm = 0
        self.j = 0
        self.last = None

    def begin(self, starts, epsilon):
        rp = self.replay
        self.idx.copy_(starts)
        self.frames.copy_(self.trainer.preprocess_state(rp['obs'][self.idx], per_sample=False))
        self.eps.fill_(float(epsilon))
        self.losses.zero_()
        if self.trainer.config.stack_internal_states:
            self.trainer.model.init_internal_states(self.B)
        self.j = 0

    def load_batch(self, new_obs, action, reward, value):
        """Non-resident mode (host-fed): copy this step's batch into the static input buffers."""
        self.inp['new_obs'].copy_(new_obs, non_blocking=True)
        self.inp['action'].copy_(action, non_blocking=True)
        self.inp['reward'].copy_(reward, non_blocking=True)
        self.inp['value'].copy_(value, non_blocking=True)

    def _body(self):
        tr, rp, B = self.trainer, self.replay, self.B
        if self.resident:  # this step's batch out of the device-resident replay: one gather launch into static buffers
            from . import _lib
            no, ac, rw, va = rp['new_obs'], rp['action'], rp['reward'], rp['value']
            ok = (no.dtype == torch.uint8 and ac.dtype == torch.float32 and rw.dtype == torch.int64
                  and va.dtype == torch.float32 and all(t.is_contiguous() for t in (no, ac, rw, va)))
            if not ok:
                raise TypeError('StepRunner: replay tensors must be contiguous u8 / f32 / i64 / f32 (DeviceReplay)')
            _lib.check(_lib.lib().sb_gather_step(self.idx.data_ptr(), B, no.shape[0], no.data_ptr(),
                                                 no[0].numel(), ac.data_ptr(), ac.shape[1], rw.data_ptr(), va.data_ptr(),
                                                 self.inp['new_obs'].data_ptr(), self.inp['action'].data_ptr(),
                                                 self.inp['reward'].data_ptr(), self.inp['value'].data_ptr(),
                                                 _lib.stream_ptr()), 'sb_gather_step')
        new_states, action, reward, value = (self.inp['new_obs'], self.inp['action'], self.inp['reward'],
                                             self.inp['value'])
        # scheduled sampling (trainer.py:125-132): ground truth w.p. epsilon, else own argmax decode; input noise with
        # the per-frame median; frame-stack shift -- one kernel, in place on the static frame buffer.  Issued right
        # after the backward pass (nothing later reads the frames): under data parallelism it hides behind the last
        # gradient all-reduce
        def feedback(out):
            ops.frame_feedback(self.frames, new_states, out['argmax'], self.eps, float(tr.config.input_noise),
                               tr.model._seed_dev)
        fused_log = len(self.names) == 5  # sb_loss_total adds (total, reconstruct, value, reward, lstm) to the log itself
        out = tr.train_step(self.frames, action, new_states, reward, value, self.eps, after_backward=feedback,
                            loss_acc=self.losses if fused_log else None)
        if not fused_log:
            self.losses += torch.stack([out[k].detach().float() for k in self.names])
        self.idx += 1
        self.last = {k: v.detach() for k, v in out.items()}  # no reference to this step's autograd graph survives

    def step(self):
        opt = self.trainer.optimizer
        model = self.trainer.model
        if self.graph is not None and self.model_key != model.capture_key():
            self.graph, self.warm = None, 0  # a buffer the graph baked in was re-allocated: capture again
        if self.j == 0 or not self.use_graph:
            self._body()
        elif self.graph is None and self.warm < 2:
            self._body()  # eager warm-up of the steady-state step before capture
            self.warm += 1
        elif self.graph is None:
            if os.environ.get('SB_BENCH_DEBUG'):
                import sys
                print('[StepRunner] capturing the steady-state step', file=sys.stderr, flush=True)
            torch.cuda.synchronize()
            opt.zero_grad(set_to_none=True)
            from . import _lib
            g = torch.cuda.CUDAGraph()
            l0 = _lib.lib().sb_launch_count()
            with torch.cuda.graph(g):
                self._body()
            self.graph_kernels = int(_lib.lib().sb_launch_count() - l0)
            self.graph, self.graph_key, self.model_key = g, opt.last_key, model.capture_key()
            g.replay()  # capture only records: run the step that was just captured
            self.replays += 1
            if os.environ.get('SB_BENCH_DEBUG'):
                torch.cuda.synchronize()
                print('[StepRunner] captured and replayed once', file=sys.stderr, flush=True)
        else:
            opt.advance_for_replay(self.graph_key)
            self.graph.replay()
            self.replays += 1
        self.j += 1
