"""B200-native NextFramePredictor (drop-in for simple/next_frame_predictor.py of the reference).

Same constructor, same 120 parameters / state_dict keys and tensor layouts, same forward signature:
    forward(x [B,12,H,W] f32 in [0,1], action [B,n] or [B,1,n] one-hot, target [B,3,H,W] or None, epsilon)
        -> (logits [B,256,3,H,W], reward_pred [B,3], value_pred [B])
The arithmetic runs in the CUDA kernels of simple_b200/csrc (tcgen05 tap GEMMs + fused elementwise stages) on
NHWC bf16 activations; parameters stay fp32 in the reference layout and are re-packed to bf16 GEMM operands.
Only the default architecture flags are supported (no multi-backend dispatch); anything else raises.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import latent, ops
from .ops import ConvSpec, PackedWeight, PrepSpec, fused_prep, plain, s2d_in, s2d_out, tap_conv

# channels of the gate / embedding operand XS = [recurrent state 64 | x_start 12 | zero pad]: 76 rounded up to the tcgen05
# K step (16).  The second 64-channel TMA box of a pixel is zero-filled past channel 80 and only its first K=16 step runs.
XS_C = 80


# ---------------------------------------------------------------------------------------------------------------
# parameter containers (same attribute names and creation order as the reference => identical init + state_dict)
# ---------------------------------------------------------------------------------------------------------------
class ActionInjector(nn.Module):  # simple/utils.py:91-96
    def __init__(self, n_action, size):
        super().__init__()
        self.dense1 = nn.Linear(n_action, size)
        self.dense2 = nn.Linear(n_action, size)

    # (all seven injectors are evaluated at once by latent.HeadsIn; the broadcast multiply-add runs inside sb_prep)


class MeanAttention(nn.Module):  # simple/utils.py:72-88
    def __init__(self, in_size, out_size):
        super().__init__()
        self.input_embedding = nn.Conv2d(in_size, 4, 1)
        self.dense = nn.Linear(5 * in_size, out_size)

    def forward(self, x):  # x: contiguous NCHW fp32
        b, c = x.shape[0], x.shape[1]
        m = x.mean(dim=(2, 3))
        a = F.conv2d(x, self.input_embedding.weight, self.input_embedding.bias).contiguous()
        s = F.softmax(a.reshape(b, -1, 4), dim=1)  # reference quirk: groups are (flat index) mod 4
        s = s.reshape(b, 1, 4, x.shape[2], x.shape[3])
        am = (x.unsqueeze(2) * s).mean(dim=(3, 4))
        l = torch.cat((am, m.unsqueeze(-1)), dim=-1).reshape(b, 5 * c)
        return self.dense(l)


class RewardEstimator(nn.Module):  # next_frame_predictor.py:11-24
    def __init__(self, config, input_size):
        super().__init__()
        self.dense1 = nn.Linear(input_size, 128)
        self.dense2 = nn.Linear(128, 3)

    def forward(self, x):
        return self.dense2(F.relu(self.dense1(x)))


class ValueEstimator(nn.Module):  # :27-34
    def __init__(self, input_size):
        super().__init__()
        self.dense = nn.Linear(input_size, 1)


class MiddleNetwork(nn.Module):  # :37-50
    def __init__(self, config, filters):
        super().__init__()
        layers = []
        for i in range(config.hidden_layers):
            layers.append(nn.Conv2d(filters, filters, 3, padding=1))
            layers.append(None if i == 0 else nn.InstanceNorm2d(filters, affine=True, eps=1e-6))
        self.middle_network = nn.ModuleList(layers)


class BitsPredictor(nn.Module):  # :66-121
    def __init__(self, config, input_size, state_size, total_number_bits, bits_at_once=8):
        super().__init__()
        self.config = config
        self.total_number_bits = total_number_bits
        self.bits_at_once = bits_at_once
        self.dense1 = nn.Linear(input_size, state_size)
        self.dense2 = nn.Linear(input_size, state_size)
        self.dense3 = nn.Linear(input_size, state_size)
        self.dense4 = nn.Linear(2 ** bits_at_once, state_size)
        self.dense5 = nn.Linear(state_size, 2 ** bits_at_once)
        self.lstm = nn.LSTMCell(state_size, state_size)
        self._sampler_tables = None
        self.cache_sampler_tables = False  # set while the weights are known to be static (NextFramePredictor.repack)

    def refresh_sampler_tables(self):
        """(Re)compute the sampling kernel's weight-only operands INTO the persistent table tensors (same addresses)."""
        self._sampler_tables = ops.lstm_sampler_tables(self.lstm, self.dense5, self.dense4, out=self._sampler_tables)

    def forward(self, x, rand, target_bits=None):
        """Sampling pass on a flattened [B,4608] (C,H,W order) bottleneck: bits in {-1,+1} (stand-alone surface; the
        model itself runs both passes of the predictor inside latent.LatentTrain / latent.latent_eval)."""
        if target_bits is not None:
            raise NotImplementedError('the teacher-forced pass is part of simple_b200.latent.LatentTrain')
        n = self.total_number_bits // self.bits_at_once
        B = x.shape[0]
        w123 = torch.cat((self.dense1.weight, self.dense2.weight, self.dense3.weight), dim=0)
        b123 = torch.cat((self.dense1.bias, self.dense2.bias, self.dense3.bias), dim=0)
        first, h, c = ops.dense(x.float(), w123, b_trans=True, bias=b123).chunk(3, dim=1)
        noise = rand.sample_noise(B, n) if hasattr(rand, 'sample_noise') else None
        # weight-only operands.  While the weights are static (rollouts) they live in PERSISTENT tensors that
        # `refresh_sampler_tables` rewrites in place: a captured rollout graph keeps reading valid, current memory
        if self.cache_sampler_tables and self._sampler_tables is not None:
            tables = self._sampler_tables
        else:
            tables = ops.lstm_sampler_tables(self.lstm, self.dense5, self.dense4)
        bits, _ = ops.lstm_sample(first, h, c, self.lstm, self.dense5, self.dense4, noise=noise,
                                  seed=getattr(rand, 'seed', None) if noise is None else None, tables=tables)
        return bits, 0.0


class StochasticModel(nn.Module):  # :124-206
    def __init__(self, config, layer_shape, n_action):
        super().__init__()
        self.config = config
        channels = config.frame_shape[0] * (int(config.stacking) + 1)
        filters = [128, 512]
        self.lstm_loss = None
        self.input_embedding = nn.Conv2d(channels, config.hidden_size, 1)
        self.conv1 = nn.Conv2d(config.hidden_size, filters[0], 8, 4, padding=2)
        self.conv2 = nn.Conv2d(filters[0], filters[1], 8, 4, padding=2)
        self.dense1 = nn.Linear(2 * 2 * channels, config.bottleneck_bits)
        self.dense2 = nn.Linear(config.bottleneck_bits, layer_shape[0])
        self.dense3 = nn.Linear(config.bottleneck_bits, layer_shape[0])
        self.action_injector = ActionInjector(n_action, config.hidden_size)
        self.mean_attentions = nn.ModuleList([MeanAttention(f, 2 * channels) for f in filters])
        self.bits_predictor = BitsPredictor(config, layer_shape[0] * layer_shape[1] * layer_shape[2],
                                            config.latent_state_size, config.bottleneck_bits)

    def get_lstm_loss(self, reset=True):
        """Accumulated teacher-forcing loss since the last reset (next_frame_predictor.py:202-206 of the reference).
        The accumulator is `None` while empty: no device tensor outlives a step, so a captured CUDA graph never bakes
        in the address of a tensor that eager code later frees (round 1 kept a zeros() tensor here; every replay of
        the step graph then added whatever had since been allocated at that address to the reported loss)."""
        res = self.lstm_loss
        if res is None:
            res = torch.zeros((), device=self.dense1.weight.device)
        if reset:
            self.lstm_loss = None
        return res


def timing_signal_sep(C, H, W):
    """Separable form of the timing signal consumed by the fused kernels: [(H+W), C/2] fp32.  Rows 0..H-1 hold the
    first C/2 channels (functions of the row index only), rows H..H+W-1 the last C/2 (functions of the column)."""
    t = timing_signal_nhwc(C, H, W)
    return torch.cat((t[:, 0, :C // 2], t[0, :, C // 2:]), dim=0).contiguous()


def timing_signal_nhwc(C, H, W, min_timescale=1.0, max_timescale=1.0e4):
    """simple/utils.py:129-154 as an NHWC fp32 table [H,W,C] (channels: sin(h w), cos(h w), sin(w w), cos(w w))."""
    nts = C // 4
    inc = math.log(float(max_timescale) / float(min_timescale)) / (nts - 1)
    inv = min_timescale * torch.exp(torch.arange(nts, dtype=torch.float32) * -torch.tensor(inc, dtype=torch.float32))
    t = torch.zeros(H, W, C)
    ph = torch.arange(H, dtype=torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    pw = torch.arange(W, dtype=torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    t[:, :, 0 * nts:1 * nts] = torch.sin(ph)[:, None, :]
    t[:, :, 1 * nts:2 * nts] = torch.cos(ph)[:, None, :]
    t[:, :, 2 * nts:3 * nts] = torch.sin(pw)[None, :, :]
    t[:, :, 3 * nts:4 * nts] = torch.cos(pw)[None, :, :]
    return t.contiguous()


class DeviceRand:
    """Fast-path randomness: every draw of the forward pass is generated INSIDE the kernels from a counter-based hash of
    (seed, stream id, element index) -- the dropout masks in the fused elementwise stages, the latent noise / mix masks
    in csrc/latent.cu, Philox in the LSTM sampler -- and regenerated by the backward kernels.  Returning None tells the
    operators to do that; `seed` is an int64 DEVICE tensor so that CUDA-graph replays draw fresh numbers."""

    def __init__(self, device, seed):
        self.device = device
        self.seed = seed

    def keep_mask(self, shape, p):
        return None

    def big_keep_mask(self, shape, p):
        return None

    def uniform(self, shape):
        return None

    def truncnorm(self, shape):
        return None


class InjectedRand:
    """Parity mode ("randomness as input"): replays an explicit list of draws recorded from the oracle, in the
    reference's draw order.  Keep masks arrive as NCHW 0/1 floats."""

    def __init__(self, log, device):
        self.log = [(k, t.to(device)) for k, t in log]
        self.pos = 0
        self.device = device
        self.seed = 0
        # Optional test hook: discrete decisions (posterior pre-activations, sampled latent symbols) recorded from the
        # oracle.  bf16 upstream error can flip a near-tie; forcing the recorded decision keeps the rest of the
        # comparison meaningful while the natural flip rate is reported in `mismatch`.
        self.force = {}
        self.mismatch = {}

    def _next(self, kind, shape):
        k, t = self.log[self.pos]
        self.pos += 1
        assert k == kind and tuple(t.shape) == tuple(shape), (k, kind, tuple(t.shape), tuple(shape))
        return t

    def keep_mask(self, shape, p):
        return self._next('keep', shape)

    def big_keep_mask(self, shape, p):
        m = self._next('keep', shape)  # [B,C,H,W] -> u8 NHWC
        return m.permute(0, 2, 3, 1).contiguous().to(torch.uint8)

    def uniform(self, shape):
        return self._next('uniform', shape)

    def exponential(self, shape):
        return self._next('exp', shape)

    def sample_noise(self, B, n):
        """The n consecutive Exp(1) draws [B,256] of one sampling pass (one per torch.multinomial call)."""
        return torch.stack([self._next('exp', (B, 256)) for _ in range(n)], dim=1).contiguous()

    def truncnorm(self, shape):
        return self._next('truncnorm', shape)


class NextFramePredictor(nn.Module):

    def __init__(self, config, n_action):
        super().__init__()
        self.config = config
        c = config
        if (tuple(c.frame_shape) != (3, 105, 80) or int(c.stacking) != 4 or c.hidden_size != 96 or c.compress_steps != 5
                or c.filter_double_steps != 3 or c.hidden_layers != 2 or c.recurrent_state_size != 64
                or c.bottleneck_bits != 128 or c.latent_state_size != 128 or not c.stack_internal_states
                or not c.use_stochastic_model):
            raise ValueError('simple_b200.NextFramePredictor supports the default architecture flags only')
        self.n_action = n_action
        filters = c.hidden_size
        channels = c.frame_shape[0] * int(c.stacking) + c.recurrent_state_size
        self.gate = nn.Conv2d(channels, 2 * c.recurrent_state_size, 3, padding=1)
        self.input_embedding = nn.Conv2d(channels, c.hidden_size, 1)
        nn.init.normal_(self.input_embedding.bias, std=0.01)
        down = []
        shape = [filters, c.frame_shape[1], c.frame_shape[2]]
        self._enc_shapes = [list(shape)]
        for i in range(c.compress_steps):
            in_filters = filters
            if i < c.filter_double_steps:
                filters *= 2
            shape = [filters, shape[1] // 2, shape[2] // 2]
            self._enc_shapes.append(list(shape))
            down.append(nn.Conv2d(in_filters, filters, 4, stride=2, padding=1))
            down.append(nn.InstanceNorm2d(filters, affine=True, eps=1e-6))
        self.downscale_layers = nn.ModuleList(down)
        middle_shape = shape
        up = []
        inj = [ActionInjector(n_action, filters)]
        self._dec_shapes = []
        for i in range(c.compress_steps):
            inj.append(ActionInjector(n_action, filters))
            in_filters = filters
            if i >= c.compress_steps - c.filter_double_steps:
                filters //= 2
            tgt = self._enc_shapes[-i - 2]
            op = (tgt[1] - shape[1] * 2, tgt[2] - shape[2] * 2)
            assert op[0] in (0, 1) and op[1] in (0, 1)
            shape = [filters, tgt[1], tgt[2]]
            self._dec_shapes.append(list(shape))
            up.append(nn.ConvTranspose2d(in_filters, filters, 4, stride=2, padding=1, output_padding=op))
            up.append(nn.InstanceNorm2d(filters, affine=True, eps=1e-6))
        self.upscale_layers = nn.ModuleList(up)
        self.action_injectors = nn.ModuleList(inj)
        self.logits = nn.Conv2d(c.hidden_size, 256 * c.frame_shape[0], 1)
        self.middle_network = MiddleNetwork(c, middle_shape[0])
        self.reward_estimator = RewardEstimator(c, middle_shape[0] + filters)
        self.value_estimator = ValueEstimator(middle_shape[0] * middle_shape[1] * middle_shape[2])
        self.stochastic_model = StochasticModel(c, middle_shape, n_action)

        # Recurrent state lives in PERSISTENT buffers (stable addresses => CUDA-graph replayable); a forward leaves its
        # new state in `_pending` and `commit_state()` copies it in (after the backward pass, which still needs the old
        # gate operand).  A forward auto-commits a pending state first, so plain `model(x, a)` loops behave like the
        # reference.
        # One buffer set PER BATCH SIZE, never freed (`_bufs`): CUDA graphs captured for the trainer (B = batch_size) and
        # for the simulated env (B = agents) both keep valid addresses when the two alternate on one model.
        self._bufs = {}           # B -> dict(state, xs_prev, xs_buf, ps)
        self._state = None        # fp32 [B,H,W,64] recurrent state (detached)
        self._xs_prev = None      # bf16 [B,H,W,80]: previous step's (state | x_start | 0) operand of the gate conv
        self._xs_buf = None       # bf16 [B,H,W,80]: this step's operand (persistent: the zero padding is written once)
        self._ps_buf = None       # bf16 [B,H,W,64]: posterior input (x_start 12 | target 3 | zero pad)
        # bumped whenever memory that a captured graph may have baked in is re-allocated (device move, operand rebuild):
        # StepRunner / SimulatedVecEnv compare it with the value they captured under and re-capture on mismatch
        self._generation = 0
        self._has_prev = False
        self._pending = None
        self._seed_dev = None
        self._plan = None
        self._packed = None
        self._auto_repack = True
        self.parity_rand = None   # set to an InjectedRand to run with explicit randomness
        self._head_targets = None
        self._logits_mode, self._ce_target, self._ce_clip = 'tensor', None, 0.0
        self._lp = None
        self.last_aux = {}

    # ------------------------------------------------------------------ reference surface
    def init_internal_states(self, batch_size):
        dev = self.logits.weight.device
        H, W = self.config.frame_shape[1:]
        bufs = self._bufs.get(batch_size)
        if bufs is None or bufs['state'].device != dev:
            bufs = self._bufs[batch_size] = dict(
                state=torch.zeros(batch_size, H, W, self.config.recurrent_state_size, device=dev),
                xs_prev=torch.zeros(batch_size, H, W, XS_C, device=dev, dtype=torch.bfloat16),
                xs_buf=torch.zeros(batch_size, H, W, XS_C, device=dev, dtype=torch.bfloat16),
                ps=torch.zeros(batch_size, H, W, 64, device=dev, dtype=torch.bfloat16))  # channels 15..63 stay zero
        else:
            bufs['state'].zero_()  # keep the buffers (and their addresses)
        self._state, self._xs_prev, self._xs_buf, self._ps_buf = (bufs['state'], bufs['xs_prev'], bufs['xs_buf'],
                                                                  bufs['ps'])
        self._has_prev = False
        self._pending = None

    def capture_key(self, rollout=False):
        """Identity of every buffer a captured CUDA graph of this model bakes in.  A holder of such a graph compares the
        key it captured under with the current one and re-captures on mismatch (stale-graph hazards).
        `rollout`: the graph also reads the persistent sampler tables (eval forward with static weights)."""
        t = self.stochastic_model.bits_predictor._sampler_tables if rollout else None
        ptr = lambda x: None if x is None else x.data_ptr()
        return (self._generation, ptr(self._state), ptr(self._xs_prev), ptr(self._xs_buf), ptr(self._ps_buf),
                ptr(self._seed_dev), None if t is None else t['E'].data_ptr())

    def commit_state(self):
        """Make the state computed by the last forward current (next_frame_predictor.py:303,312 of the reference)."""
        if self._pending is None:
            return
        self._xs_prev.copy_(self._pending)
        self._has_prev = True
        self._pending = None

    @property
    def internal_states(self):
        self.commit_state()
        return None if self._state is None else self._state.permute(0, 3, 1, 2)

    @property
    def last_x_start(self):
        self.commit_state()
        return self._xs_prev[..., 64:76].float().permute(0, 3, 1, 2) if self._has_prev else None

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        # same shapes, new values: re-pack INTO the existing bf16 operands / sampler tables (addresses survive, so
        # captured graphs stay valid)
        if self._packed is not None and self.logits.weight.is_cuda:
            for pw in self._packed.values():
                pw.fresh = False
            bp = self.stochastic_model.bits_predictor
            self.repack(static=bp.cache_sampler_tables)
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        # device / dtype move: every derived buffer is rebuilt at new addresses
        self._packed = None
        self._plan = None
        self._bufs = {}
        self._state = self._xs_prev = self._xs_buf = self._ps_buf = None
        self._has_prev, self._pending, self._seed_dev = False, None, None
        self.stochastic_model.bits_predictor._sampler_tables = None
        self._generation += 1
        return r

    # ------------------------------------------------------------------ static plan (layouts, specs, timing tables)
    def _build_plan(self, dev):
        c = self.config
        H, W = c.frame_shape[1:]
        P = {}
        taps22 = [(dy, dx) for dy in range(2) for dx in range(2)]
        taps33 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
        # timing tables (next_frame_predictor.py:234,241,270): index 0, encoder 1..5, decoder 6..10
        es, ds = self._enc_shapes, self._dec_shapes
        tshapes = [es[0]] + [es[i] for i in range(5)] + [ds[i] for i in range(5)]
        P['timing'] = [timing_signal_sep(s[0], s[1], s[2]).to(dev) for s in tshapes]
        P['stoch_timing'] = timing_signal_sep(c.hidden_size, H, W).to(dev)
        # gate 3x3 on XS (state 0..63 | x_start 64..75 | zero pad to XS_C = 80), no dgrad (inputs are detached / data)
        HW = H * W
        P['gate'] = ConvSpec('gate', H, W, XS_C, H, W, 128, taps33, relu=False, need_dgrad=False,
                             flops=2.0 * HW * 128 * 76 * 9)
        P['embed'] = ConvSpec('embed', H, W, XS_C, H, W, 96, [(0, 0)], relu=False, flops=2.0 * HW * 96 * 76)
        P['down'], P['down_lay'] = [], []
        for i in range(5):
            Ci, Hi, Wi = es[i]
            Co, Ho, Wo = es[i + 1]
            lay = s2d_in(Hi, Wi, Ci, 4, 2, 1)
            P['down_lay'].append(lay)
            P['down'].append(ConvSpec(f'down{i}', lay.H2, lay.W2, 4 * Ci, Ho, Wo, Co, taps22, relu=True,
                                      out_f32=(Ho * Wo <= 30), flops=2.0 * Ho * Wo * Co * Ci * 16))
        P['mid'] = [ConvSpec(f'mid{i}', 3, 2, 768, 3, 2, 768, taps33, relu=True, bwd_mask=(i == 0), out_f32=True,
                             flops=2.0 * 6 * 768 * 768 * 9) for i in range(2)]
        P['up'], P['up_lay'] = [], []
        shape = es[5]
        for i in range(5):
            Ci, Hi, Wi = shape
            Co, Ho, Wo = ds[i]
            lay = s2d_out(Hi, Wi, Ho, Wo, Co)
            P['up_lay'].append(lay)
            P['up'].append(ConvSpec(f'up{i}', Hi, Wi, Ci, Hi + 1, Wi + 1, 4 * Co, [(-dy, -dx) for dy, dx in taps22],
                                    relu=True, transposed=True, bias_tile=4, out_f32=(Ho * Wo <= 30),
                                    flops=2.0 * Hi * Wi * Ci * Co * 16))
            shape = ds[i]
        P['logits'] = ConvSpec('logits', H, W, 96, H, W, 768, [(0, 0)], relu=False, flops=2.0 * HW * 768 * 96)
        P['logits'].dbias_src = self.logits_dbias = {}
        # posterior encoder of the stochastic model (train only)
        P['s_embed'] = ConvSpec('s_embed', H, W, 64, H, W, 96, [(0, 0)], relu=False, need_dgrad=False,
                                flops=2.0 * HW * 96 * 15)
        l1 = s2d_in(H, W, 96, 8, 4, 2)
        P['s_lay1'] = l1
        P['s_conv1'] = ConvSpec('s_conv1', l1.H2, l1.W2, 16 * 96, 26, 20, 128, taps22, relu=False,
                                flops=2.0 * 520 * 128 * 96 * 64)
        l2 = s2d_in(26, 20, 128, 8, 4, 2)
        P['s_lay2'] = l2
        P['s_conv2'] = ConvSpec('s_conv2', l2.H2, l2.W2, 16 * 128, 6, 5, 512, taps22, relu=False,
                                flops=2.0 * 30 * 512 * 128 * 64)
        # fused elementwise stages
        T = P['timing']
        drop = c.dropout
        P['prep_enc'] = []
        for i in range(5):
            Ci, Hi, Wi = es[i]
            P['prep_enc'].append(PrepSpec(Hi, Wi, Ci, plain(Hi, Wi, Ci), P['down_lay'][i], relu_in=(i > 0), norm=(i > 0),
                                          Ta=T[0] if i == 0 else None, want_out1=True, p=drop, stream_id=1 + i,
                                          Tb=T[i + 1]))
        P['prep_x5'] = PrepSpec(3, 2, 768, plain(3, 2, 768), plain(3, 2, 768), relu_in=True, norm=True)
        rd = c.residual_dropout
        P['prep_m0'] = PrepSpec(3, 2, 768, plain(3, 2, 768), plain(3, 2, 768), p=rd, stream_id=10)
        P['prep_m1'] = PrepSpec(3, 2, 768, plain(3, 2, 768), plain(3, 2, 768), relu_in=True, p=rd, stream_id=11)
        P['prep_dec'] = [PrepSpec(3, 2, 768, plain(3, 2, 768), plain(3, 2, 768), relu_in=True, norm=True, p=drop,
                                  stream_id=20, inject=True)]
        for i in range(1, 5):
            Co, Ho, Wo = ds[i - 1]
            P['prep_dec'].append(PrepSpec(Ho, Wo, Co, P['up_lay'][i - 1], plain(Ho, Wo, Co), relu_in=True, norm=True,
                                          Ta=T[5 + i], p=drop, stream_id=20 + i, inject=True))
        P['prep_final'] = PrepSpec(H, W, 96, P['up_lay'][4], plain(H, W, 96), relu_in=True, norm=True, Ta=T[10])
        P['prep_s0'] = PrepSpec(H, W, 96, plain(H, W, 96), l1, Ta=P['stoch_timing'], inject=True)
        P['prep_s1'] = PrepSpec(26, 20, 128, plain(26, 20, 128), l2, relu_in=True)
        P['tmean10'] = timing_signal_nhwc(96, H, W).mean(dim=(0, 1)).to(dev)
        # convs whose output has exactly one consumer, a fused prep stage: that stage's backward emits their bias grad
        for sp in [P['embed'], P['s_embed'], P['mid'][1]] + P['down'] + P['up']:
            sp.dbias_src = {}
        return P

    def _build_packed(self):
        c = self.config
        sm = self.stochastic_model
        pk = {}
        # gate input order is (state 64, x_start 12) == XS order; pad K 76 -> XS_C (a multiple of the MMA K step)
        pk['gate'] = PackedWeight(self.gate.weight, 1, Ipad=XS_C)
        # input_embedding input order is (x_start 12, state 64): permute to XS order
        perm = [64 + i for i in range(12)] + list(range(64))
        pk['embed'] = PackedWeight(self.input_embedding.weight, 1, Ipad=XS_C, iperm=perm)
        for i in range(5):
            pk[f'down{i}'] = PackedWeight(self.downscale_layers[2 * i].weight, 2)
            pk[f'up{i}'] = PackedWeight(self.upscale_layers[2 * i].weight, 2)
        for i in range(2):
            pk[f'mid{i}'] = PackedWeight(self.middle_network.middle_network[2 * i].weight, 1)
        # logits: output channel v*3+colour -> colour*256+v so that a softmax group is contiguous
        operm = [(o % 3) * 256 + o // 3 for o in range(768)]
        pk['logits'] = PackedWeight(self.logits.weight, 1, operm=operm)
        pk['s_embed'] = PackedWeight(sm.input_embedding.weight, 1, Ipad=64)
        pk['s_conv1'] = PackedWeight(sm.conv1.weight, 4)
        pk['s_conv2'] = PackedWeight(sm.conv2.weight, 4)
        self._logit_perm = torch.tensor(operm, device=self.logits.weight.device)
        return pk

    def _weights(self):
        return {'gate': self.gate, 'embed': self.input_embedding, 'logits': self.logits,
                's_embed': self.stochastic_model.input_embedding, 's_conv1': self.stochastic_model.conv1,
                's_conv2': self.stochastic_model.conv2,
                **{f'down{i}': self.downscale_layers[2 * i] for i in range(5)},
                **{f'up{i}': self.upscale_layers[2 * i] for i in range(5)},
                **{f'mid{i}': self.middle_network.middle_network[2 * i] for i in range(2)}}

    def repack(self, names=None, static=False):
        """Refresh the bf16 GEMM operands from the fp32 master weights (after an optimiser step / load)."""
        bp = self.stochastic_model.bits_predictor
        bp.cache_sampler_tables = bool(static)  # static=True: the caller promises frozen weights until the next repack
        if static:
            bp.refresh_sampler_tables()
        if self._packed is None:
            self._packed = self._build_packed()
            self._generation += 1
            return
        mods = self._weights()
        for k in (names or self._packed.keys()):
            self._packed[k].pack(mods[k].weight)

    def packed_weight_map(self, defer_unpack=None):
        """{conv weight parameter: its PackedWeight}; `defer_unpack` switches packed-gradient mode (see ops.PackedWeight)
        on or off for all of them.  Valid after the first forward (the operands are built lazily)."""
        if self._packed is None:
            return {}
        mods = self._weights()
        if defer_unpack is not None:
            for pw in self._packed.values():
                pw.defer_unpack = bool(defer_unpack)
                pw.gpacked = None
                pw.grad_slot = None  # (a data-parallel gradient arena re-attaches its slots right after, every step)
                pw.on_ready = None
        return {mods[k].weight: pw for k, pw in self._packed.items()}

    def _latent_params(self):
        """Flat parameter groups of the latent / heads block (rebuilt after a device move)."""
        lp = getattr(self, '_lp', None)
        if lp is None or not lp.valid():
            lp = self._lp = latent.LatentParams(self)
        return lp

    def _conv(self, name, spec, a, need_weight_grad=True):
        m = self._weights_cache[name]
        return tap_conv(spec, a, m.weight, m.bias, self._packed[name])

    # ------------------------------------------------------------------ forward
    def _trunk(self, x, action, target, epsilon, rand):
        c = self.config
        P = self._plan
        dev = x.device
        B = x.shape[0]
        H, W = c.frame_shape[1:]
        seed = rand.seed
        training = self.training and target is not None
        aux = self.last_aux = {}

        def keep(spec_shape_nchw, p):
            return rand.big_keep_mask(spec_shape_nchw, p)

        # ---- standardise + recurrent state (K5, K6)
        if self._state is None or self._state.shape[0] != B or self._state.device != dev:
            self.init_internal_states(B)
        # XS = [state 64 | x_start 12 | zero pad to 80] lives in a persistent buffer whose padding is zeroed once (a fresh
        # torch.zeros of 137 MB per step was pure fill traffic); a detached alias keeps this step's autograd graph off it
        xs = self._xs_buf.detach()
        if not self._has_prev:
            xs[..., :64].zero_()  # first step of a window: the recurrent state is zero and the gate conv is skipped
        # x_start goes into the gate / embedding operand and, in the training branch, into the posterior's operand
        ps = self._ps_buf.detach() if target is not None else None  # channels 15..63 stay zero (init_internal_states)
        ops.standardize(x.contiguous().float(), dst=xs, coff=64, dst2=ps if training else None, coff2=0)
        if self._has_prev:
            G = self._conv('gate', P['gate'], self._xs_prev)
            xs = ops.GruBlend.apply(G, self._state, xs, self._xs_prev)  # fp32 state updated in place
        self._pending = xs.detach()
        e = self._conv('embed', P['embed'], xs)
        # ---- encoder (K1, K3, K4, K8)
        skips = []
        main = e
        for i in range(5):
            sp = P['prep_enc'][i]
            nrm = self.downscale_layers[2 * i - 1] if i > 0 else None
            km = keep((B, sp.C, sp.H, sp.W), sp.p)
            skip, A = fused_prep(sp, main, gamma=nrm.weight if nrm is not None else None,
                                 beta=nrm.bias if nrm is not None else None, seed=seed, keep=km,
                                 dbias_dst=(P['embed'] if i == 0 else P['down'][i - 1]).dbias_src)
            skips.append(skip)
            main = self._conv(f'down{i}', P['down'][i], A)
        nrm = self.downscale_layers[9]
        _, x5 = fused_prep(P['prep_x5'], main, gamma=nrm.weight, beta=nrm.bias, dbias_dst=P['down'][4].dbias_src)
        # ---- bottleneck heads and latent (csrc/latent.cu around plain GEMMs; flatten order of the reference: C,H,W)
        lp = self._latent_params()
        outs = latent.HeadsIn.apply(x5, action, lp)
        value_pred, h_flat = outs[0], outs[1]
        inj = [(outs[2 + 2 * i], outs[3 + 2 * i]) for i in range(6)]  # decoder injectors 1..5, then the posterior's
        if target is not None:
            ops.standardize(target, dst=ps, coff=12, out_f32=target)  # in place, like the reference (:332-334)
        sm = self.stochastic_model
        if training:
            x12 = self._posterior_features(ps, inj[5])
            hm, x_mid, lstm_loss = latent.LatentTrain.apply(h_flat, x12, lp, epsilon, rand, c, aux)
            sm.lstm_loss = lstm_loss if sm.lstm_loss is None else sm.lstm_loss + lstm_loss
        else:
            hm, x_mid = latent.latent_eval(h_flat.detach(), lp, rand, aux)
        # ---- middle network (:52-63)
        mn = self.middle_network.middle_network
        sp = P['prep_m0']
        _, A = fused_prep(sp, hm, seed=seed, keep=keep((B, 768, 3, 2), sp.p))
        y1 = self._conv('mid0', P['mid'][0], A)
        sp = P['prep_m1']
        _, A = fused_prep(sp, y1, seed=seed, keep=keep((B, 768, 3, 2), sp.p))
        y2 = self._conv('mid1', P['mid'][1], A)
        # ---- decoder (K2, K3, K4, K7, K8)
        main, add, nrm = y2, y1, mn[3]
        for i in range(5):
            sp = P['prep_dec'][i]
            sc, sh = inj[i]
            km = keep((B, sp.C, sp.H, sp.W), sp.p)
            _, A = fused_prep(sp, main, add=add, gamma=nrm.weight, beta=nrm.bias, sc=sc, sh=sh, seed=seed, keep=km,
                              dbias_dst=(P['mid'][1] if i == 0 else P['up'][i - 1]).dbias_src)
            main = self._conv(f'up{i}', P['up'][i], A)
            add = skips[4 - i]
            nrm = self.upscale_layers[2 * i + 1]
        _, hf = fused_prep(P['prep_final'], main, add=add, gamma=nrm.weight, beta=nrm.bias,
                           dbias_dst=P['up'][4].dbias_src)
        # x_fin = mean_hw(IN(z)*gamma + beta + timing) == beta + mean_hw(timing): the normalised part has zero mean
        x_fin = nrm.bias + P['tmean10']
        tg = self._head_targets
        if tg is not None or not torch.is_grad_enabled():
            # fused tail: reward MLP (+ reward CE, value MSE and all their gradients when the targets are known)
            reward_pred, l_rew, l_val = latent.HeadsOut.apply(x_mid, x_fin, value_pred, tg[0] if tg else None,
                                                              tg[1] if tg else None, self)
            aux['loss_reward'], aux['loss_value'] = l_rew, l_val
        else:  # plain `model(x, a, target)` under autograd with the losses taken outside: differentiable library ops
            reward_pred = self.reward_estimator(torch.cat((x_mid, x_fin.unsqueeze(0).expand(B, -1)), dim=1))
        if self._logits_mode == 'tensor':
            logits = self._conv('logits', P['logits'], hf)  # bf16 [B,H,W,768], channel = colour*256 + value
        else:  # logits GEMM fused with the softmax-CE / arg-max decode: the logits never reach HBM (ops.LogitsCE)
            m, pw, logits = self._weights_cache['logits'], self._packed['logits'], None
            if self._logits_mode == 'ce':
                aux['loss_reconstruct'], aux['argmax'] = ops.LogitsCE.apply(P['logits'], hf, m.weight, m.bias, pw,
                                                                             self._ce_target, self._ce_clip)
            elif self._logits_mode == 'ce_raw':  # the raw f64 sum; trainer.TotalLoss finishes all losses in one launch
                aux['rec_sum'], aux['argmax'] = ops.LogitsCE.apply(P['logits'], hf, m.weight, m.bias, pw,
                                                                   self._ce_target, self._ce_clip, True)
                aux['rec_n'] = float(B * 3 * hf.shape[1] * hf.shape[2])
            else:
                aux['argmax'] = ops.logits_argmax(P['logits'], hf, m.bias, pw)
        return logits, reward_pred, value_pred

    def _posterior_features(self, ps, inj_s):
        """Convolutional part of StochasticModel.forward's training branch (next_frame_predictor.py:163-174): the two
        8x8/4 convs and their MeanAttention read-outs -> [B,60]."""
        P = self._plan
        sm = self.stochastic_model
        e = self._conv('s_embed', P['s_embed'], ps)
        sc, sh = inj_s
        _, A = fused_prep(P['prep_s0'], e, sc=sc, sh=sh, dbias_dst=P['s_embed'].dbias_src)
        c1 = self._conv('s_conv1', P['s_conv1'], A)  # [B,26,20,128], pre-ReLU
        x1 = ops.mean_attention(c1, sm.mean_attentions[0])
        _, A = fused_prep(P['prep_s1'], c1)
        c2 = self._conv('s_conv2', P['s_conv2'], A)  # [B,6,5,512]
        x2 = ops.mean_attention(c2, sm.mean_attentions[1])
        return torch.cat((x1, x2), dim=-1)

    def forward_internal(self, x, action, target=None, epsilon=0, rewards=None, values=None, logits_mode='tensor',
                         ce_target=None, ce_clip=0.0):
        """Runs the network and returns the INTERNAL logits (bf16 [B,H,W,768], colour-major) + reward + value.
        `rewards` (long [B]) / `values` (f32 [B]): when given, the reward CE and value MSE of simple/trainer.py:138-141
        are produced by the fused heads kernel and left in `last_aux['loss_reward' / 'loss_value']`.
        `logits_mode`: 'tensor' materialises the logits; 'ce' fuses the logits GEMM with the clipped pixel CE against
        `ce_target` (u8 [B,3,H,W]) -> `last_aux['loss_reconstruct' / 'argmax']`, logits = None; 'argmax' (rollouts)
        leaves only the decoded frame in `last_aux['argmax']`."""
        self._head_targets = (rewards, values) if rewards is not None else None
        self._logits_mode, self._ce_target, self._ce_clip = logits_mode, ce_target, ce_clip
        dev = x.device
        if dev.type != 'cuda':
            raise RuntimeError('simple_b200.NextFramePredictor runs on CUDA (sm_100a) only; there is no CPU fallback')
        if self._plan is None:
            self._plan = self._build_plan(dev)
        if self._packed is None:
            self._packed = self._build_packed()
            self._generation += 1
        elif self._auto_repack:
            self.repack()
        self._weights_cache = self._weights()
        self.logits_dbias.clear()
        self.commit_state()
        if self._seed_dev is None or self._seed_dev.device != dev:
            # derived from torch.manual_seed() without consuming the global generator
            self._seed_dev = torch.full((1,), torch.initial_seed() & ((1 << 40) - 1), dtype=torch.int64, device=dev)
        if self.parity_rand is not None:
            rand = self.parity_rand
        else:
            self._seed_dev.add_(1)  # new Philox streams every forward; the backward of this step re-reads the value
            rand = DeviceRand(dev, self._seed_dev)
        return self._trunk(x, action, target, epsilon, rand)

    def forward(self, x, action, target=None, epsilon=0):
        logits, reward_pred, value_pred = self.forward_internal(x, action, target, epsilon)
        B, H, W, _ = logits.shape
        ref = logits.reshape(B, H, W, 3, 256).permute(0, 4, 3, 1, 2).float()
        return ref, reward_pred, value_pred
