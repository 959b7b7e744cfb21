"""CPU oracle for the SimPLe world-model hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 PyTorch-on-CPU *restatement* of the reference algorithm
(thomas-schillaci/SimPLe).  It is a checker: only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product path (`simple_b200`) never does.

Parity pinning: the reference ships no tests and no golden vectors, so
this restatement is pinned against the reference itself, imported live from
/root/reference in the build container by `oracle/make_golden.py`, which also writes
the fixtures under `tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle
against those fixtures wherever the tests run.

Strict-validation mode (`cfg.emulate_bf16 = True`, see `bf16_config`): the same restatement, but rounded to bf16
at exactly the points where the CUDA path stores bf16 (GEMM operands and outputs, the skip tensors, the outputs of
the fused elementwise stages; DESIGN.md section 5 lists them) with fp32 everywhere the CUDA path keeps fp32 (tiny
deep-layer outputs, statistics, the latent path).  It is still a CPU fp32 computation -- what remains different from
the GPU is summation order and intrinsic precision -- so discrete decisions (posterior signs, sampled latent symbols,
arg-max frames) can be compared END TO END instead of stage by stage.  With the flag absent / False every function
below is the pinned fp32 restatement, bit-identical to the reference.

All randomness is explicit: every function takes a `rand` object (see `GlobalRand`,
`ReplayRand`).  `GlobalRand` draws from the torch / numpy global generators with exactly
the primitive calls and in exactly the order the reference uses, so with equal seeds the
oracle and the reference agree bit for bit on CPU; it records every draw so that the
same tensors can be replayed into the CUDA path ("randomness as input").

Citations `file:line` are relative to /root/reference/.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------
# configuration (simple/__main__.py:154-191 defaults)
# ----------------------------------------------------------------------------------
def default_config(**overrides):
    cfg = dict(
        agents=16, batch_size=4, bottleneck_bits=128, bottleneck_noise=0.1, clip_grad_norm=1.0,
        compress_steps=5, device='cpu', done_on_last_rollout_step=True, dropout=0.15,
        env_name='Freeway', epochs=15, experiment_name='x', filter_double_steps=3,
        frame_shape=(3, 105, 80), hidden_layers=2, hidden_size=96, input_noise=0.05,
        latent_rnn_max_sampling=0.5, latent_state_size=128, latent_use_max_probability=0.8,
        load_models=False, noop_max=8, ppo_gamma=0.99, ppo_lr=1e-4, recurrent_state_size=64,
        render_evaluation=False, render_training=False, residual_dropout=0.5, rollout_length=50,
        save_models=False, scheduled_sampling_decay_steps=22250,
        simulation_flip_first_random_for_beginning=True, stacking=4, stack_internal_states=True,
        target_loss_clipping=0.03, use_ppo_lr_decay=False, use_stochastic_model=True, use_wandb=False,
    )
    cfg.update(overrides)
    return SimpleNamespace(**cfg)


def bf16_config(**overrides):
    """default_config with the bf16-emulation (strict validation) mode switched on."""
    return default_config(emulate_bf16=True, **overrides)


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 (nearest even) and back; straight-through gradient (the CUDA path rounds gradients too, but
    unbiasedly: the oracle keeps them in fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def _emul(cfg):
    return bool(getattr(cfg, 'emulate_bf16', False))


def _q(cfg, x):
    """A point where the CUDA path stores bf16 (identity in the pinned fp32 mode)."""
    return _RoundBF16.apply(x) if _emul(cfg) else x


# ----------------------------------------------------------------------------------
# randomness sources
# ----------------------------------------------------------------------------------
class GlobalRand:
    """Draws from the global torch/numpy RNGs exactly like the reference and records them."""

    def __init__(self):
        self.log = []

    def _rec(self, kind, t):
        self.log.append((kind, t.clone()))
        return t

    def keep_mask(self, shape, p):
        # F.dropout(x, p) == x * empty_like(x).bernoulli_(1 - p) / (1 - p)  (ATen native dropout)
        return self._rec('keep', torch.empty(shape).bernoulli_(1 - p))

    def uniform(self, shape):
        # torch.rand_like (simple/utils.py:161, next_frame_predictor.py:184)
        return self._rec('uniform', torch.rand(shape))

    def exponential(self, shape):
        # torch.multinomial(w, 1) == argmax(w / empty_like(w).exponential_(1))  (ATen multinomial, n_sample=1)
        return self._rec('exp', torch.empty(shape).exponential_(1))

    def truncnorm(self, shape):
        # scipy.stats.truncnorm.rvs(-2, 2, size, scale=0.2) on numpy's global RNG (next_frame_predictor.py:179)
        from scipy.stats import truncnorm
        t = torch.tensor(truncnorm.rvs(-2, 2, size=tuple(shape), scale=0.2), dtype=torch.float32)
        return self._rec('truncnorm', t)


class ReplayRand:
    """Replays a recorded list of draws (same order)."""

    def __init__(self, log):
        self.log = list(log)
        self.pos = 0

    def _next(self, kind, shape):
        k, t = self.log[self.pos]
        self.pos += 1
        assert k == kind, (k, kind)
        assert tuple(t.shape) == tuple(shape), (kind, tuple(t.shape), tuple(shape))
        return t

    def keep_mask(self, shape, p):
        return self._next('keep', shape)

    def uniform(self, shape):
        return self._next('uniform', shape)

    def exponential(self, shape):
        return self._next('exp', shape)

    def truncnorm(self, shape):
        return self._next('truncnorm', shape)


# ----------------------------------------------------------------------------------
# helpers (simple/utils.py)
# ----------------------------------------------------------------------------------
def standardize_frame(x):
    """simple/utils.py:122-126.  x: [..., C, H, W]; per channel mean / UNBIASED var over H*W."""
    mean = x.mean(dim=(-1, -2), keepdim=True)
    var = x.var(dim=(-1, -2), keepdim=True)  # unbiased
    n = torch.tensor(float(x.shape[-1] * x.shape[-2]))
    return (x - mean) / torch.max(torch.sqrt(var), torch.rsqrt(n))


def timing_signal_nd(shape, min_timescale=1.0, max_timescale=1.0e4):
    """simple/utils.py:129-154.  shape = (C, H, W) -> [1, C, H, W]."""
    channels = shape[0]
    nts = channels // 4
    inc = torch.log(torch.tensor(float(max_timescale) / float(min_timescale))) / \
        (torch.tensor(nts, dtype=torch.float32) - 1)
    inv = min_timescale * torch.exp(torch.arange(nts, dtype=torch.float32) * -inc)
    res = torch.zeros(tuple(shape))
    for dim in range(2):
        length = shape[dim + 1]
        pos = torch.arange(length, dtype=torch.float32)
        st = pos.unsqueeze(1) * inv.unsqueeze(0)
        sig = torch.cat((torch.sin(st), torch.cos(st)), dim=1)
        prepad = dim * 2 * nts
        postpad = channels - (dim + 1) * 2 * nts
        sig = F.pad(sig, [prepad, postpad]).permute(1, 0)
        nd = [1, channels, 1, 1]
        nd[dim + 2] = -1
        res = res + sig.reshape(nd)
    return res


def mix(x1, x2, eps, rand):
    """simple/utils.py:157-163: take x2 where rand < eps."""
    mask = (rand.uniform(x1.shape) < eps).float()
    return (1 - mask) * x1 + mask * x2


def action_injector(p, prefix, x, action):
    """simple/utils.py:98-105."""
    c = x.shape[1]
    m = F.linear(action, p[prefix + '.dense1.weight'], p[prefix + '.dense1.bias']).reshape(-1, c, 1, 1)
    x = x * torch.sigmoid(m)
    m = F.linear(action, p[prefix + '.dense2.weight'], p[prefix + '.dense2.bias']).reshape(-1, c, 1, 1)
    return x + m


def mean_attention(p, prefix, x):
    """simple/utils.py:79-88.  Note the memory reinterpretation `a.view(B, -1, 4)`:
    softmax groups are (flat index over (4,H,W)) mod 4."""
    b = x.shape[0]
    m = x.mean(dim=(2, 3))
    a = F.conv2d(x, p[prefix + '.input_embedding.weight'], p[prefix + '.input_embedding.bias'])
    s = a.reshape(b, -1, 4)
    s = F.softmax(s, dim=1)
    s = s.reshape(b, 1, 4, *x.shape[2:])
    am = (x.unsqueeze(2) * s).mean(dim=(3, 4))
    l = torch.cat((am, m.unsqueeze(-1)), dim=-1).reshape(b, 5 * x.shape[1])
    return F.linear(l, p[prefix + '.dense.weight'], p[prefix + '.dense.bias'])


def bit_to_int(bits01):
    """simple/utils.py:108-112, LSB first.  bits01: [..., 8] of {0,1} -> int64 [...]."""
    w = (2 ** torch.arange(bits01.shape[-1])).to(bits01.dtype)
    return (bits01 * w).sum(-1).long()


def int_to_bit(ints, nbits=8):
    """simple/utils.py:115-119, LSB first."""
    x = ints.unsqueeze(-1).int()
    return torch.cat([torch.remainder(x // 2 ** i, 2) for i in range(nbits)], -1).float()


class _InstanceNormEmul(torch.autograd.Function):
    """InstanceNorm as the fused CUDA stages compute it (csrc/elementwise.cu): statistics single-pass (E[z^2] - E[z]^2,
    clamped at 0) for the streaming kernel, two-pass for planes of <= 64 pixels (stats_small_kernel); backward by the
    analytic formula rstd * gamma * (g - mean(g) - xhat * mean(g * xhat))."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        mean = x.mean(dim=(2, 3), keepdim=True)
        if x.shape[2] * x.shape[3] <= 64:
            var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
        else:
            var = ((x * x).mean(dim=(2, 3), keepdim=True) - mean * mean).clamp(min=0.)
        rstd = 1.0 / torch.sqrt(var + eps)
        xhat = (x - mean) * rstd
        ctx.save_for_backward(xhat, rstd, w)
        return xhat * w.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)

    @staticmethod
    def backward(ctx, g):
        xhat, rstd, w = ctx.saved_tensors
        gx = g * w.reshape(1, -1, 1, 1)
        m1 = gx.mean(dim=(2, 3), keepdim=True)
        m2 = (gx * xhat).mean(dim=(2, 3), keepdim=True)
        return rstd * (gx - m1 - xhat * m2), (g * xhat).sum(dim=(0, 2, 3)), g.sum(dim=(0, 2, 3)), None


def instance_norm(x, w, b, eps=1e-6, cfg=None):
    if cfg is not None and _emul(cfg):
        return _InstanceNormEmul.apply(x, w, b, eps)
    return F.instance_norm(x, weight=w, bias=b, eps=eps)


def lstm_cell(p, prefix, x, h, c):
    gates = F.linear(x, p[prefix + '.weight_ih'], p[prefix + '.bias_ih']) + \
        F.linear(h, p[prefix + '.weight_hh'], p[prefix + '.bias_hh'])
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


# ----------------------------------------------------------------------------------
# BitsPredictor / StochasticModel (simple/next_frame_predictor.py:66-206)
# ----------------------------------------------------------------------------------
def bits_predictor(p, cfg, x, rand, target_bits=None, prefix='stochastic_model.bits_predictor', aux=None):
    nsteps = cfg.bottleneck_bits // 8
    x = torch.flatten(x, start_dim=1)
    first = F.linear(x, p[prefix + '.dense1.weight'], p[prefix + '.dense1.bias'])
    h = F.linear(x, p[prefix + '.dense2.weight'], p[prefix + '.dense2.bias'])
    c = F.linear(x, p[prefix + '.dense3.weight'], p[prefix + '.dense3.bias'])
    if target_bits is not None:  # teacher forced, :87-107
        tb = target_bits.reshape(-1, nsteps, 8).clamp(min=0.)
        tints = bit_to_int(tb)
        thot = F.one_hot(tints, 256).float()
        emb = F.linear(thot, p[prefix + '.dense4.weight'], p[prefix + '.dense4.bias'])
        emb = emb * (rand.keep_mask(emb.shape, 0.1) / 0.9)
        tin = torch.cat((first.unsqueeze(1), emb), dim=1)
        outs = []
        for i in range(nsteps):
            h, c = lstm_cell(p, prefix + '.lstm', tin[:, i, :], h, c)
            outs.append(h)
        outs = torch.stack(outs, dim=1)
        outs = outs * (rand.keep_mask(outs.shape, 0.1) / 0.9)
        pred = F.linear(outs, p[prefix + '.dense5.weight'], p[prefix + '.dense5.bias'])
        loss = F.cross_entropy(pred.permute(0, 2, 1), tints)
        return pred, loss / cfg.rollout_length
    outs = []
    inp = first
    for i in range(nsteps):  # sampling, :109-121
        h, c = lstm_cell(p, prefix + '.lstm', inp, h, c)
        logits = F.linear(h, p[prefix + '.dense5.weight'], p[prefix + '.dense5.bias'])
        w = torch.exp(logits / 1.0)  # atari_utils/atari_utils/utils.py:95-101
        noisy = w / rand.exponential(w.shape)
        sample = torch.argmax(noisy, dim=-1)
        if aux is not None:  # log-domain scores of the categorical draw: how close a decision was (tests: near-ties)
            aux.setdefault('sample_scores', []).append(torch.log(noisy).detach())
        outs.append(sample)
        inp = F.linear(F.one_hot(sample, 256).float(), p[prefix + '.dense4.weight'], p[prefix + '.dense4.bias'])
    outs = torch.stack(outs, dim=1)
    if aux is not None:
        aux['sample_scores'] = torch.stack(aux['sample_scores'], dim=1)  # [B, 16, 256]
        aux['samples'] = outs
    bits = int_to_bit(outs).reshape(-1, cfg.bottleneck_bits)
    return 2 * bits - 1, 0.0


def add_bits(p, layer, bits, prefix='stochastic_model'):
    """:153-159"""
    z_mul = torch.sigmoid(F.linear(bits, p[prefix + '.dense2.weight'], p[prefix + '.dense2.bias']))
    z_add = F.linear(bits, p[prefix + '.dense3.weight'], p[prefix + '.dense3.bias'])
    return layer * z_mul[:, :, None, None] + z_add[:, :, None, None]


def stochastic_model(p, cfg, layer, inputs, action, target, epsilon, rand, training, aux):
    """:161-200.  Returns (x, lstm_loss_increment)."""
    pre = 'stochastic_model'
    if training and target is not None:
        q = lambda t: _q(cfg, t)
        x = q(torch.cat((inputs, target), dim=1))  # bf16 operand (x_start | standardised target)
        x = q(F.conv2d(x, q(p[pre + '.input_embedding.weight']), p[pre + '.input_embedding.bias']))
        x = x + timing_signal_nd((cfg.hidden_size, *cfg.frame_shape[1:]))
        x = q(action_injector(p, pre + '.action_injector', x, action))
        x = q(F.conv2d(x, q(p[pre + '.conv1.weight']), p[pre + '.conv1.bias'], stride=4, padding=2))
        x1 = mean_attention(p, pre + '.mean_attentions.0', x)
        x = F.relu(x)
        x = q(F.conv2d(x, q(p[pre + '.conv2.weight']), p[pre + '.conv2.bias'], stride=4, padding=2))
        x2 = mean_attention(p, pre + '.mean_attentions.1', x)
        x = torch.cat((x1, x2), dim=-1)
        x = F.linear(x, p[pre + '.dense1.weight'], p[pre + '.dense1.bias'])
        aux['posterior_pre'] = x
        bits_clean = (2 * (0 < x).float()).detach() - 1
        x = x + rand.truncnorm(x.shape)
        x = torch.tanh(x)
        bits = x + (2 * (0 < x).float() - 1 - x).detach()
        noise = rand.uniform(x.shape)
        noise = 2 * (cfg.bottleneck_noise < noise).float() - 1
        bits = bits * noise
        bits = mix(bits, x, 1 - epsilon, rand)
        _, lstm_loss = bits_predictor(p, cfg, layer, rand, bits_clean)
        bits_pred, _ = bits_predictor(p, cfg, layer, rand, aux=aux)
        aux['bits_pred'] = bits_pred
        bits_pred = bits_clean + (bits_pred - bits_clean).detach()
        bits = mix(bits_pred, bits, 1 - (1 - epsilon) * cfg.latent_rnn_max_sampling, rand)
        aux['bits'] = bits
        res = add_bits(p, layer, bits)
        return mix(res, layer, 1 - (1 - epsilon) * cfg.latent_use_max_probability, rand), lstm_loss
    bits, _ = bits_predictor(p, cfg, layer, rand, aux=aux)
    aux['bits'] = bits
    return add_bits(p, layer, bits), 0.0


# ----------------------------------------------------------------------------------
# NextFramePredictor.forward (simple/next_frame_predictor.py:286-359)
# ----------------------------------------------------------------------------------
def layer_shapes(cfg):
    """Shapes of the 11 main timing signals and the convT output paddings (:232-270)."""
    filters = cfg.hidden_size
    shape = [filters, *cfg.frame_shape[1:]]
    tshapes = [list(shape)]
    shapes = [list(shape)]
    for i in range(cfg.compress_steps):
        if i < cfg.filter_double_steps:
            filters *= 2
        tshapes.append(list(shape))
        shape = [filters, shape[1] // 2, shape[2] // 2]
        shapes.append(list(shape))
    opads = []
    for i in range(cfg.compress_steps):
        if i >= cfg.compress_steps - cfg.filter_double_steps:
            filters //= 2
        shape = [filters, shape[1] * 2, shape[2] * 2]
        op = (0 if shape[1] == shapes[-i - 2][1] else 1, 0 if shape[2] == shapes[-i - 2][2] else 1)
        shape = [filters, shape[1] + op[0], shape[2] + op[1]]
        opads.append(op)
        tshapes.append(list(shape))
    return tshapes, opads


def init_state(cfg, batch):
    """init_internal_states :286-290"""
    return {'internal_states': torch.zeros(batch, cfg.recurrent_state_size, *cfg.frame_shape[1:]),
            'last_x_start': None}


def forward(p, cfg, state, x, action, target=None, epsilon=0, rand=None, training=True, aux=None):
    """One NextFramePredictor.forward.  `state` is mutated like the module's attributes.
    Returns (logits [B,256,3,H,W], reward_pred [B,3], value_pred [B], lstm_loss_increment)."""
    if aux is None:
        aux = {}
    tshapes, opads = layer_shapes(cfg)
    timing = [timing_signal_nd(s) for s in tshapes]
    q = lambda t: _q(cfg, t)  # bf16 storage points of the CUDA path (identity unless cfg.emulate_bf16)
    x_start = torch.stack([standardize_frame(f) for f in x])  # :307
    # get_internal_states :292-304
    internal = state['internal_states']
    if state['last_x_start'] is not None:
        sa = torch.cat((q(internal), q(state['last_x_start'])), dim=1)  # gate operand: bf16 copy of the fp32 state
        gc = q(F.conv2d(sa, q(p['gate.weight']), p['gate.bias'], padding=1))
        g, c = torch.split(gc, cfg.recurrent_state_size, dim=1)
        g = torch.sigmoid(g)
        c = torch.tanh(c)
        internal = internal * g
        internal = internal + c * (1 - g)
        state['internal_states'] = internal.detach()
    h = torch.cat((q(x_start), q(internal)), dim=1)
    state['last_x_start'] = x_start
    h = q(F.conv2d(h, q(p['input_embedding.weight']), p['input_embedding.bias'])) + timing[0]
    inputs = []
    for i in range(cfg.compress_steps):  # :320-326
        inputs.append(q(h))  # skip tensor (stored bf16)
        h = h * (rand.keep_mask(h.shape, cfg.dropout) / (1 - cfg.dropout))
        h = q(h + timing[i + 1])
        h = F.conv2d(h, q(p[f'downscale_layers.{2 * i}.weight']), p[f'downscale_layers.{2 * i}.bias'],
                     stride=2, padding=1)
        h = F.relu(h)
        if h.shape[2] * h.shape[3] > 30:
            h = q(h)  # conv outputs are bf16 except the tiny deep layers (<= 30 pixels), which stay fp32
        h = instance_norm(h, p[f'downscale_layers.{2 * i + 1}.weight'], p[f'downscale_layers.{2 * i + 1}.bias'], cfg=cfg)
    h = q(h)  # the bottleneck is stored bf16
    aux['skips'] = inputs
    aux['x5'] = h
    value_pred = F.linear(torch.flatten(h, start_dim=1), p['value_estimator.dense.weight'],
                          p['value_estimator.dense.bias']).squeeze(-1)
    h = action_injector(p, 'action_injectors.0', h, action)
    if target is not None:  # :332-334 (in place on the caller's tensor)
        for b in range(len(target)):
            target[b] = standardize_frame(target[b])
    h, lstm_loss = stochastic_model(p, cfg, h, x_start, action, target, epsilon, rand, training, aux)
    aux['after_stoch'] = h
    x_mid = h.mean(dim=(2, 3))
    h = q(h)
    # MiddleNetwork :52-63
    for i in range(cfg.hidden_layers):
        y = q(h * (rand.keep_mask(h.shape, cfg.residual_dropout) / (1 - cfg.residual_dropout)))
        y = F.relu(F.conv2d(y, q(p[f'middle_network.middle_network.{2 * i}.weight']),
                            p[f'middle_network.middle_network.{2 * i}.bias'], padding=1))
        if i == 0:
            h = y
        else:
            h = instance_norm(h + y, p[f'middle_network.middle_network.{2 * i + 1}.weight'],
                              p[f'middle_network.middle_network.{2 * i + 1}.bias'], cfg=cfg)
    inputs = list(reversed(inputs))
    for i in range(cfg.compress_steps):  # :343-350
        h = h * (rand.keep_mask(h.shape, cfg.dropout) / (1 - cfg.dropout))
        h = q(action_injector(p, f'action_injectors.{i + 1}', h, action))
        h = F.conv_transpose2d(h, q(p[f'upscale_layers.{2 * i}.weight']), p[f'upscale_layers.{2 * i}.bias'],
                               stride=2, padding=1, output_padding=opads[i])
        h = F.relu(h)
        if h.shape[2] * h.shape[3] > 30:
            h = q(h)
        h = h + inputs[i]
        h = instance_norm(h, p[f'upscale_layers.{2 * i + 1}.weight'], p[f'upscale_layers.{2 * i + 1}.bias'], cfg=cfg)
        h = h + timing[1 + cfg.compress_steps + i]
    x_fin = h.mean(dim=(2, 3))
    r = torch.cat((x_mid, x_fin), dim=1)
    r = F.relu(F.linear(r, p['reward_estimator.dense1.weight'], p['reward_estimator.dense1.bias']))
    reward_pred = F.linear(r, p['reward_estimator.dense2.weight'], p['reward_estimator.dense2.bias'])
    h = q(h)
    aux['final_x'] = h
    logits = q(F.conv2d(h, q(p['logits.weight']), p['logits.bias']))
    logits = logits.reshape(-1, 256, *cfg.frame_shape)
    return logits, reward_pred, value_pred, lstm_loss


# ----------------------------------------------------------------------------------
# losses (simple/trainer.py:134-144)
# ----------------------------------------------------------------------------------
def wm_losses(cfg, logits, reward_pred, value_pred, lstm_loss, new_states, rewards, values):
    ce = F.cross_entropy(logits, new_states, reduction='none')
    ce = torch.max(ce, torch.tensor(cfg.target_loss_clipping))
    loss_reconstruct = ce.mean() - cfg.target_loss_clipping
    loss_value = F.mse_loss(value_pred, values)
    loss_reward = F.cross_entropy(reward_pred, rewards)
    loss = loss_reconstruct + loss_value + loss_reward + lstm_loss
    return loss, loss_reconstruct, loss_value, loss_reward


def pixel_ce_np(logits, targets, clip):
    """numpy float64 restatement of the per-pixel 256-way CE (+ argmax, first-index ties),
    on logits laid out [N, 256]; used to check the fused CUDA kernel.  trainer.py:134-137,130."""
    lg = np.asarray(logits, dtype=np.float64)
    m = lg.max(axis=1, keepdims=True)
    e = np.exp(lg - m)
    s = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    ce = lse - lg[np.arange(lg.shape[0]), targets]
    active = ce > clip
    soft = e / s
    onehot = np.zeros_like(soft)
    onehot[np.arange(lg.shape[0]), targets] = 1.0
    dlogits = (soft - onehot) * active[:, None] / lg.shape[0]
    loss = np.maximum(ce, clip).mean() - clip
    return loss, dlogits, lg.argmax(axis=1)


# ----------------------------------------------------------------------------------
# clip_grad_norm_ + Adafactor.step (simple/trainer.py:148-149; simple/adafactor.py:113-212)
# ----------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ semantics: g *= clamp(max_norm / (||g|| + 1e-6), max=1)."""
    total = torch.norm(torch.stack([torch.norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def adafactor_step(param, grad, st):
    """One parameter, fairseq defaults (adafactor.py:45-57).  `st` is the state dict (mutated)."""
    eps1, eps2, clip_threshold, decay_rate = 1e-30, 1e-3, 1.0, -0.8
    factored = grad.dim() >= 2
    if len(st) == 0:
        st['step'] = 0
        if factored:
            st['exp_avg_sq_row'] = torch.zeros(grad.shape[:-1])
            st['exp_avg_sq_col'] = torch.zeros(grad.shape[:-2] + grad.shape[-1:])
        else:
            st['exp_avg_sq'] = torch.zeros_like(grad)
        st['RMS'] = 0
    st['step'] += 1
    st['RMS'] = param.norm(2) / (param.numel() ** 0.5)
    lr = max(eps2, float(st['RMS'])) * min(1e-2, 1.0 / math.sqrt(st['step']))
    beta2t = 1.0 - math.pow(st['step'], decay_rate)
    update = grad ** 2 + eps1
    if factored:
        row, col = st['exp_avg_sq_row'], st['exp_avg_sq_col']
        row.mul_(beta2t).add_(update.mean(dim=-1), alpha=1.0 - beta2t)
        col.mul_(beta2t).add_(update.mean(dim=-2), alpha=1.0 - beta2t)
        r = (row / row.mean(dim=-1, keepdim=True)).rsqrt_().unsqueeze(-1)
        c = col.unsqueeze(-2).rsqrt()
        update = torch.mul(r, c) * grad
    else:
        v = st['exp_avg_sq']
        v.mul_(beta2t).add_(update, alpha=1.0 - beta2t)
        update = v.rsqrt() * grad
    update = update / (update.norm(2) / (update.numel() ** 0.5) / clip_threshold).clamp_(min=1.0)
    update = update * lr
    return param - update


# ----------------------------------------------------------------------------------
# simulated rollout step (simple/subproc_vec_env.py:409-445) -- arithmetic only
# ----------------------------------------------------------------------------------
def rollout_step(p, cfg, state, frames, actions_onehot, rand, last_step):
    """frames: f32 [A,12,H,W] holding k/255.  Returns (new_frames f32, obs_u8 [A,12,H,W], reward f64 [A])."""
    with torch.no_grad():
        logits, rewards, values, _ = forward(p, cfg, state, frames, actions_onehot, None, 0, rand, training=False)
    new_states = torch.argmax(logits, dim=1)
    frames = torch.cat((frames[:, 3:], new_states.float() / 255), dim=1)
    obs = (frames * 255).byte()
    rew = (torch.argmax(rewards, dim=1) - 1).numpy().astype('float')
    if last_step:
        rew = rew + values.numpy().astype('float')
    return frames, obs, rew


# ----------------------------------------------------------------------------------
# Trainer.preprocess_state median (simple/trainer.py:58-65)
# ----------------------------------------------------------------------------------
def input_noise(state_u8, keep_mask):
    """state/255; where keep_mask==0 replace by torch.median of the whole tensor (lower median)."""
    s = state_u8.float() / 255
    return s * keep_mask + torch.median(s) * (1 - keep_mask)
