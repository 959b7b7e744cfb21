"""World-model trainer (drop-in for simple/trainer.py of the reference): Trainer(model, config).train(epoch, env, steps).

Same loop semantics (windows of `rollout_length` consecutive transitions, input noise with median replacement,
scheduled sampling in epoch 0, four/five losses, clip_grad_norm_, Adafactor, one optimiser step per time step) but
device-resident: the replay list is mirrored once into contiguous device tensors, windows are gathered on the device,
losses are accumulated on the device and read back once per window, and clip+Adafactor are fused CUDA kernels.
"""
import os

import torch
import torch.nn.functional as F

from . import ops
from .adafactor import Adafactor


def wm_loss(model, config, frames, actions, new_states_u8, rewards, values, epsilon, keep_logits=False):
    """Forward + the losses of simple/trainer.py:134-144.  Returns a dict of device scalars (no host sync).

    frames f32 [B,12,H,W] in [0,1]; actions f32 [B,n] one-hot; new_states_u8 u8 [B,3,H,W]; rewards long [B];
    values f32 [B].  The pixel cross-entropy, its gradient and the argmax decode are ONE fused kernel pass over the
    bf16 logits (ops.PixelCE); `argmax` is the u8 frame the reference obtains with torch.argmax (trainer.py:130).
    """
    target = new_states_u8.float() / 255
    logits, reward_pred, value_pred = model.forward_internal(frames, actions, target, epsilon)
    out = {}
    if keep_logits:  # tests only: the fused loss overwrites the logits with their gradient
        B, H, W, _ = logits.shape
        out['logits_ref'] = logits.detach().reshape(B, H, W, 3, 256).permute(0, 4, 3, 1, 2).float().clone()
    loss_reconstruct, argmax = ops.PixelCE.apply(logits, new_states_u8.contiguous(), float(config.target_loss_clipping),
                                                 True, True)
    loss_value = F.mse_loss(value_pred, values)
    loss_reward = F.cross_entropy(reward_pred, rewards)
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
            from atari_utils.logger import WandBLogger  # optional dependency of the reference
            self.logger = WandBLogger()
        self.optimizer = Adafactor(self.model.parameters())
        self.model_step = 1
        self.reward_step = 1
        self._replay = None
        self._replay_key = None
        self.grad_sync = None  # set by simple_b200.parallel for data-parallel training

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
            med = torch.median(state)
        return state * keep + med * (1 - keep)

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

    def train_step(self, frames, actions, new_states_u8, rewards, values, epsilon):
        """One optimiser step (trainer.py:122-149).  Returns the loss dict of `wm_loss` (device scalars)."""
        self.model.train()
        out = wm_loss(self.model, self.config, frames, actions, new_states_u8, rewards, values, epsilon)
        self.optimizer.zero_grad(set_to_none=True)
        out['loss'].backward()
        grads = self.grad_sync(self.model) if self.grad_sync is not None else None
        self.optimizer.step(max_grad_norm=self.config.clip_grad_norm, grads=grads)
        return out

    def train(self, epoch, env, steps=15000):
        if epoch == 0:
            steps *= 3
        cfg = self.config
        c, h, w = cfg.frame_shape
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
        rp = self._replay
        try:
            from tqdm import trange
            iterator = trange(0, steps, rollout_len, desc='Training world model', unit_scale=rollout_len)
        except ImportError:
            iterator = range(0, steps, rollout_len)
        n_losses = 5 if cfg.use_stochastic_model else 4
        names = ['loss', 'loss_reconstruct', 'loss_value', 'loss_reward', 'loss_lstm'][:n_losses]
        metrics = {}
        for i in iterator:
            epsilon = self.epsilon_schedule(epoch, i)
            starts = rp.sample_starts(B)
            frames = self.preprocess_state(rp.obs[starts], per_sample=False)
            if cfg.stack_internal_states:
                self.model.init_internal_states(B)
            losses = torch.zeros(n_losses, device=dev)
            for j in range(rollout_len):
                idx = starts + j
                new_states_u8 = rp.new_obs[idx]
                out = self.train_step(frames, rp.action[idx], new_states_u8, rp.reward[idx], rp.value[idx], epsilon)
                if j < rollout_len - 1:
                    # scheduled sampling (trainer.py:125-132): ground truth w.p. epsilon, else own argmax decode
                    use_gt = (torch.rand(B, device=dev) < epsilon).reshape(B, 1, 1, 1)
                    nxt = torch.where(use_gt, new_states_u8, out['argmax'])
                    nxt = self.preprocess_state(nxt, per_sample=True)
                    frames = torch.cat((frames[:, c:], nxt), dim=1)
                losses += torch.stack([out[k].detach().float() for k in names])
            host = (losses / rollout_len).tolist()  # the only host sync of the window
            metrics = dict(zip(names, host))
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
