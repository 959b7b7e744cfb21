"""Latent / heads block of the world model: everything between the bottleneck and the middle network, the reward and
value heads and their losses, as fused CUDA kernels (csrc/latent.cu) around the skinny fp32 GEMMs of `sb_dense` (csrc/dense.cu).

Reference: simple/next_frame_predictor.py:153-200 (StochasticModel), :66-121 (BitsPredictor), :11-34 (heads), :328-339,
:352-354; simple/utils.py:91-105 (ActionInjector), :157-163 (mix); simple/trainer.py:138-141 (value / reward losses).

Three autograd Functions with HAND-WRITTEN backward passes (autograd only links them):
  HeadsIn      bottleneck x5 (bf16 NHWC) -> value_pred, injected bottleneck h (flat fp32), scale/shift of all injectors
  LatentTrain  h, MeanAttention features -> posterior bits, teacher-forced LSTM loss, sampled bits, add_bits, mixes
               -> middle-network operand (bf16 NHWC), x_mid, lstm loss       (LatentEval: the rollout branch, no grad)
  HeadsOut     x_mid, x_fin -> reward_pred, reward CE, value MSE (+ their gradients, produced in the forward launch)
Parameter gradients of the dense layers are written straight into `p.grad` (views of one flat buffer per group of layers);
the parameters of a group live in ONE contiguous buffer (`FlatBank`) so that e.g. dense1 | dense2 | dense3 of the
BitsPredictor is a single [B,4608] x [4608,384] GEMM without any concatenation.
"""
import ctypes

import torch

from . import _lib
from .ops import dense, zeros as _zeros

INJ_WIDTHS = (768, 768, 768, 768, 384, 192, 96)  # action_injectors.0-5, stochastic_model.action_injector
INJ_S = sum(INJ_WIDTHS)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _seed(rand):
    s = getattr(rand, 'seed', 0)
    if isinstance(s, torch.Tensor):
        return s.data_ptr(), 0
    return None, int(s) & 0xFFFFFFFFFFFFFFFF


def _eps(epsilon):
    if isinstance(epsilon, torch.Tensor):
        return epsilon.data_ptr(), 0.0
    return None, float(epsilon)


def _acc_grad(p, g):
    """`p.grad (+)= g` without going through autograd (g: a view of a gradient buffer that outlives the step)."""
    if p.grad is None:
        p.grad = g
    else:
        p.grad = p.grad + g


class FlatBank:
    """Several Parameters with equal trailing dims stored back to back in ONE buffer (`p.data` become views of it)."""

    def __init__(self, params):
        self.params = list(params)
        self.rows = [p.shape[0] for p in self.params]
        self.flat = torch.cat([p.detach() for p in self.params], dim=0).contiguous()
        self._bind()

    def _bind(self):
        off = 0
        for p, r in zip(self.params, self.rows):
            p.data = self.flat[off:off + r]
            off += r

    def valid(self):
        off, esz = 0, self.flat.element_size() * (self.flat[0].numel() if self.flat.dim() > 1 else 1)
        for p, r in zip(self.params, self.rows):
            if p.device != self.flat.device or p.data_ptr() != self.flat.data_ptr() + off * esz:
                return False
            off += r
        return True

    def split_grad(self, g):
        """g: gradient w.r.t. `flat` -> accumulated into the members' .grad as views."""
        off = 0
        for p, r in zip(self.params, self.rows):
            _acc_grad(p, g[off:off + r])
            off += r


class LatentParams:
    """Flat parameter groups of one NextFramePredictor (rebuilt when the parameters move to another device)."""

    def __init__(self, model):
        sm, bp = model.stochastic_model, model.stochastic_model.bits_predictor
        inj = list(model.action_injectors) + [sm.action_injector]
        assert tuple(m.dense1.weight.shape[0] for m in inj) == INJ_WIDTHS
        self.inj_w = FlatBank([m.dense1.weight for m in inj] + [m.dense2.weight for m in inj])   # [2S, n]
        self.inj_b = FlatBank([m.dense1.bias for m in inj] + [m.dense2.bias for m in inj])       # [2S]
        self.w123 = FlatBank([bp.dense1.weight, bp.dense2.weight, bp.dense3.weight])             # [384, 4608]
        self.b123 = FlatBank([bp.dense1.bias, bp.dense2.bias, bp.dense3.bias])
        self.w23 = FlatBank([sm.dense2.weight, sm.dense3.weight])                                # [1536, 128]
        self.b23 = FlatBank([sm.dense2.bias, sm.dense3.bias])
        self.widths = (ctypes.c_int * len(INJ_WIDTHS))(*INJ_WIDTHS)
        self.model = model

    def valid(self):
        return all(b.valid() for b in (self.inj_w, self.inj_b, self.w123, self.b123, self.w23, self.b23))


# ---------------------------------------------------------------------------------------------------------------
class HeadsIn(torch.autograd.Function):
    """value_pred = value_estimator(x5); h = ActionInjector_0(x5); (scale, shift) of the other six injectors."""

    @staticmethod
    def forward(ctx, x5, action, lp):
        L, st = _lib.lib(), _lib.stream_ptr()
        B, dev = x5.shape[0], x5.device
        assert x5.dtype == torch.bfloat16 and x5.is_contiguous() and x5.numel() == B * 4608
        a = action.reshape(B, -1).float()
        bank = dense(a, lp.inj_w.flat, b_trans=True, bias=lp.inj_b.flat)  # [B, 2S]: dense1 pre-activations | dense2 outputs
        ve = lp.model.value_estimator.dense
        scsh = torch.empty(2, B * INJ_S, device=dev)
        h_flat = torch.empty(B, 4608, device=dev)
        value = torch.empty(B, device=dev)
        _lib.check(L.sb_heads_fwd(x5.data_ptr(), bank.data_ptr(), lp.widths, len(INJ_WIDTHS), ve.weight.data_ptr(),
                                  ve.bias.data_ptr(), B, scsh[0].data_ptr(), scsh[1].data_ptr(), h_flat.data_ptr(),
                                  value.data_ptr(), st), 'sb_heads_fwd')
        ctx.save_for_backward(x5, bank, a)
        ctx.lp = lp
        ctx.set_materialize_grads(False)
        outs = [value, h_flat]
        off = INJ_WIDTHS[0]
        for C in INJ_WIDTHS[1:]:
            outs.append(scsh[0, off * B:(off + C) * B].view(B, C))
            outs.append(scsh[1, off * B:(off + C) * B].view(B, C))
            off += C
        return tuple(outs)

    @staticmethod
    def backward(ctx, dvalue, dh, *dscsh):
        x5, bank, a = ctx.saved_tensors
        lp = ctx.lp
        L, st = _lib.lib(), _lib.stream_ptr()
        B, dev = x5.shape[0], x5.device
        g = _lib.InjGrads()
        keep = []
        for i in range(1, len(INJ_WIDTHS)):
            for kind, t in (('sc', dscsh[2 * (i - 1)]), ('sh', dscsh[2 * (i - 1) + 1])):
                if t is None:
                    continue
                if t.dtype != torch.float32:
                    t = t.float()
                keep.append(t)
                if kind == 'sc':
                    g.dsc[i], g.sc_sb[i], g.sc_sc[i] = t.data_ptr(), t.stride(0), t.stride(1)
                else:
                    g.dsh[i], g.sh_sb[i], g.sh_sc[i] = t.data_ptr(), t.stride(0), t.stride(1)
        if dvalue is not None:
            dvalue = dvalue.contiguous().float()
        if dh is not None:
            dh = dh.contiguous()
        ve = lp.model.value_estimator.dense
        dx5 = torch.empty_like(x5)
        dbank = torch.empty(B, 2 * INJ_S, device=dev)
        small = _zeros((4608 + 1 + 2 * INJ_S,), dev)  # dwv | dbv | dbank_b
        dwv, dbv, dbank_b = small[:4608], small[4608:4609], small[4609:]
        _lib.check(L.sb_heads_bwd(x5.data_ptr(), bank.data_ptr(), lp.widths, len(INJ_WIDTHS), ctypes.byref(g),
                                  ve.weight.data_ptr(), _ptr(dvalue), _ptr(dh), B, dx5.data_ptr(), dbank.data_ptr(),
                                  dwv.data_ptr(), dbv.data_ptr(), dbank_b.data_ptr(), st), 'sb_heads_bwd')
        lp.inj_w.split_grad(dense(dbank, a, a_trans=True))
        lp.inj_b.split_grad(dbank_b)
        if dvalue is not None:
            _acc_grad(ve.weight, dwv.view(1, 4608))
            _acc_grad(ve.bias, dbv)
        return dx5, None, None


# ---------------------------------------------------------------------------------------------------------------
def _sampler_operands(bp, bsum, L, st):
    """Transposed LSTM / dense5 weights and the symbol table E of the sampling kernel, from the CURRENT weights."""
    dev = bp.lstm.weight_hh.device
    whhT = torch.empty(128, 512, device=dev)
    wihT = torch.empty(128, 512, device=dev)
    w5T = torch.empty(128, 256, device=dev)
    emb = torch.empty(256, 128, device=dev)
    _lib.check(L.sb_sampler_tables(bp.lstm.weight_hh.data_ptr(), bp.lstm.weight_ih.data_ptr(), bp.dense5.weight.data_ptr(),
                                   bp.dense4.weight.data_ptr(), bp.dense4.bias.data_ptr(), whhT.data_ptr(),
                                   wihT.data_ptr(), w5T.data_ptr(), emb.data_ptr(), st), 'sb_sampler_tables')
    E = dense(emb, wihT, bias=bsum)  # [256,512]: input-gate pre-activations of every symbol
    return whhT, wihT, w5T, E


def _sample_bits(first, h0, c0, hc_ld, bsum, wihT, whhT, w5T, E, b5, noise, seed_ptr, B, L, st):
    dev = first.device
    G0 = dense(first, wihT, bias=bsum)  # [B,512]
    bits = torch.empty(B, 128, device=dev)
    ints = torch.empty(B, 16, device=dev, dtype=torch.int32)
    _lib.check(L.sb_lstm_sample2(G0.data_ptr(), h0.data_ptr(), c0.data_ptr(), E.data_ptr(), whhT.data_ptr(),
                                 w5T.data_ptr(), b5.data_ptr(), _ptr(noise), seed_ptr if noise is None else None, B, hc_ld,
                                 bits.data_ptr(), ints.data_ptr(), st), 'sb_lstm_sample2')
    return bits


class LatentTrain(torch.autograd.Function):
    """StochasticModel.forward, training branch (next_frame_predictor.py:162-197), on the injected bottleneck h (flat) and
    the concatenated MeanAttention features x12 [B,60].  Returns (hm bf16 [B,3,2,768], x_mid [B,768], lstm_loss)."""

    @staticmethod
    def forward(ctx, h_flat, x12, lp, epsilon, rand, cfg, aux):
        L, st = _lib.lib(), _lib.stream_ptr()
        B, dev = h_flat.shape[0], h_flat.device
        sm = lp.model.stochastic_model
        bp = sm.bits_predictor
        seed_ptr, seed_val = _seed(rand)
        eps_ptr, eps_val = _eps(epsilon)
        f32 = dict(device=dev, dtype=torch.float32)
        # ---- posterior bits (draw order of the reference: truncnorm, flip noise, first mix)
        pre = dense(x12, sm.dense1.weight, b_trans=True, bias=sm.dense1.bias)  # [B,128]
        tn, u_flip, u1 = rand.truncnorm(pre.shape), rand.uniform(pre.shape), rand.uniform(pre.shape)
        x = torch.empty(B, 128, **f32)
        bits2 = torch.empty(B, 128, **f32)
        bits_clean = torch.empty(B, 128, **f32)
        flags = torch.empty(B, 128, device=dev, dtype=torch.uint8)
        tints = torch.empty(B, 16, device=dev, dtype=torch.int32)
        _lib.check(L.sb_posterior_bits_fwd(pre.data_ptr(), _ptr(tn), _ptr(u_flip), _ptr(u1), eps_ptr, eps_val,
                                           float(cfg.bottleneck_noise), seed_ptr, seed_val, B, x.data_ptr(),
                                           bits2.data_ptr(), bits_clean.data_ptr(), flags.data_ptr(), tints.data_ptr(),
                                           st), 'sb_posterior_bits_fwd')
        # ---- BitsPredictor, teacher-forced pass on the clean bits (:87-107)
        proj = dense(h_flat, lp.w123.flat, b_trans=True, bias=lp.b123.flat)  # [B,384] = first | h0 | c0
        keep1, keep2 = rand.keep_mask((B, 16, 128), 0.1), rand.keep_mask((B, 16, 128), 0.1)
        bsum = torch.empty(512, **f32)
        tin = torch.empty(B, 16, 128, **f32)
        _lib.check(L.sb_teacher_embed_fwd(proj.data_ptr(), 384, tints.data_ptr(), bp.dense4.weight.data_ptr(),
                                          bp.dense4.bias.data_ptr(), _ptr(keep1), 0.1, seed_ptr, seed_val,
                                          bp.lstm.bias_ih.data_ptr(), bp.lstm.bias_hh.data_ptr(), bsum.data_ptr(), B,
                                          tin.data_ptr(), st), 'sb_teacher_embed_fwd')
        whhT, wihT, w5T, E = _sampler_operands(bp, bsum, L, st)
        xw = dense(tin.view(B * 16, 128), wihT, bias=bsum)  # [16B,512] input-gate pre-activations of all steps
        hs = torch.empty(B, 16, 128, **f32)
        act = torch.empty(B, 16, 512, **f32)
        cs = torch.empty(B, 16, 128, **f32)
        h0, c0 = proj[:, 128:256], proj[:, 256:384]
        _lib.check(L.sb_lstm_teacher_fwd(xw.data_ptr(), h0.data_ptr(), c0.data_ptr(), whhT.data_ptr(), B, 384,
                                         hs.data_ptr(), act.data_ptr(), cs.data_ptr(), st), 'sb_lstm_teacher_fwd')
        outs_d = torch.empty(B * 16, 128, **f32)
        _lib.check(L.sb_drop_scale(hs.data_ptr(), _ptr(keep2), 0.1, seed_ptr, seed_val, 107, outs_d.data_ptr(),
                                   B * 16 * 128, st), 'sb_drop_scale')
        pred = dense(outs_d, w5T, bias=bp.dense5.bias)  # [16B,256]
        zbuf = _zeros((1 + 256 + 128 + 1536 + 128 * 256 + 128,), dev)  # loss | db5 | db1 | db23 | dw4 | db4
        loss_sum, db5, db1 = zbuf[0:1], zbuf[1:257], zbuf[257:385]
        dpred = torch.empty(B * 16, 256, **f32)
        gscale = 1.0 / (B * 16 * cfg.rollout_length)
        _lib.check(L.sb_softmax_ce_rows(pred.data_ptr(), tints.data_ptr(), B * 16, 256, gscale, loss_sum.data_ptr(),
                                        dpred.data_ptr(), db5.data_ptr(), st), 'sb_softmax_ce_rows')
        lstm_loss = (loss_sum * gscale).reshape(())
        # ---- sampling pass (:109-121; integer outputs, nothing to differentiate) and the second mix
        noise = rand.sample_noise(B, 16) if hasattr(rand, 'sample_noise') else None
        bits_pred = _sample_bits(proj[:, :128], h0, c0, 384, bsum, wihT, whhT, w5T, E, bp.dense5.bias, noise, seed_ptr, B, L, st)
        u2 = rand.uniform((B, 128))
        bits3 = torch.empty(B, 128, **f32)
        _lib.check(L.sb_mix_bits(bits_pred.data_ptr(), bits2.data_ptr(), _ptr(u2), eps_ptr, eps_val,
                                 float(cfg.latent_rnn_max_sampling), seed_ptr, seed_val, B, bits3.data_ptr(),
                                 flags.data_ptr(), st), 'sb_mix_bits')
        # ---- add_bits (:153-159) and the third mix
        zz = dense(bits3, lp.w23.flat, b_trans=True, bias=lp.b23.flat)  # [B,1536]
        u3 = rand.uniform((B, 768, 3, 2))
        hm = torch.empty(B, 3, 2, 768, device=dev, dtype=torch.bfloat16)
        x_mid = torch.empty(B, 768, **f32)
        m3 = torch.empty(B, 4608, device=dev, dtype=torch.uint8)
        _lib.check(L.sb_latent_out_fwd(h_flat.data_ptr(), zz.data_ptr(), _ptr(u3), 1, eps_ptr, eps_val,
                                       float(cfg.latent_use_max_probability), seed_ptr, seed_val, B, hm.data_ptr(),
                                       x_mid.data_ptr(), m3.data_ptr(), st), 'sb_latent_out_fwd')
        ctx.save_for_backward(h_flat, x12, x, flags, tints, proj, tin, hs, act, cs, outs_d, dpred, bits3, zz, m3, zbuf,
                              w5T, keep1, keep2)
        ctx.lp, ctx.seed = lp, (rand.seed if isinstance(getattr(rand, 'seed', 0), torch.Tensor) else None)
        ctx.seed_val = seed_val
        ctx.set_materialize_grads(False)
        if aux is not None:
            aux['posterior_pre'], aux['bits_pred'], aux['bits'] = pre, bits_pred, bits3
        return hm, x_mid, lstm_loss

    @staticmethod
    def backward(ctx, dhm, dx_mid, dlstm):
        (h_flat, x12, x, flags, tints, proj, tin, hs, act, cs, outs_d, dpred, bits3, zz, m3, zbuf, w5T, keep1,
         keep2) = ctx.saved_tensors
        lp = ctx.lp
        L, st = _lib.lib(), _lib.stream_ptr()
        B, dev = h_flat.shape[0], h_flat.device
        sm = lp.model.stochastic_model
        bp = sm.bits_predictor
        seed_ptr = ctx.seed.data_ptr() if ctx.seed is not None else None
        seed_val = ctx.seed_val
        f32 = dict(device=dev, dtype=torch.float32)
        _, db5, db1, db23 = zbuf[0:1], zbuf[1:257], zbuf[257:385], zbuf[385:1921]
        dw4, db4 = zbuf[1921:1921 + 128 * 256].view(128, 256), zbuf[1921 + 128 * 256:]
        # ---- latent out
        if dhm is not None:
            dhm = dhm.contiguous()
        dxm_ld = 0
        if dx_mid is not None:
            assert dx_mid.dtype == torch.float32 and dx_mid.stride(1) == 1
            dxm_ld = dx_mid.stride(0)
        dh_a = torch.empty(B, 4608, **f32)
        dzz = torch.empty(B, 1536, **f32)
        _lib.check(L.sb_latent_out_bwd(_ptr(dhm), _ptr(dx_mid), dxm_ld, h_flat.data_ptr(), zz.data_ptr(), m3.data_ptr(),
                                       B, dh_a.data_ptr(), dzz.data_ptr(), db23.data_ptr(), st), 'sb_latent_out_bwd')
        lp.w23.split_grad(dense(dzz, bits3, a_trans=True))
        lp.b23.split_grad(db23)
        dbits3 = dense(dzz, lp.w23.flat)  # [B,128]
        # ---- posterior bits -> dense1 of the stochastic model
        dpre = torch.empty(B, 128, **f32)
        _lib.check(L.sb_posterior_bits_bwd(dbits3.data_ptr(), x.data_ptr(), flags.data_ptr(), B, dpre.data_ptr(),
                                           db1.data_ptr(), st), 'sb_posterior_bits_bwd')
        _acc_grad(sm.dense1.weight, dense(dpre, x12, a_trans=True))
        _acc_grad(sm.dense1.bias, db1)
        dx12 = dense(dpre, sm.dense1.weight)
        # ---- teacher-forced LSTM loss (its gradient w.r.t. the logits was produced by the forward CE launch)
        _acc_grad(bp.dense5.weight, dense(dpred, outs_d, a_trans=True))
        _acc_grad(bp.dense5.bias, db5)
        douts_d = dense(dpred, bp.dense5.weight)  # [16B,128]
        dhs = torch.empty(B * 16, 128, **f32)
        _lib.check(L.sb_drop_scale(douts_d.data_ptr(), _ptr(keep2), 0.1, seed_ptr, seed_val, 107, dhs.data_ptr(),
                                   B * 16 * 128, st), 'sb_drop_scale')
        dxw = torch.empty(B * 16, 512, **f32)
        dproj = torch.empty(B, 384, **f32)
        _lib.check(L.sb_lstm_teacher_bwd(dhs.data_ptr(), act.data_ptr(), cs.data_ptr(), proj[:, 256:384].data_ptr(),
                                         bp.lstm.weight_hh.data_ptr(), B, 384, dxw.data_ptr(),
                                         dproj[:, 128:256].data_ptr(), dproj[:, 256:384].data_ptr(), st),
                   'sb_lstm_teacher_bwd')
        h_prev = torch.cat((proj[:, 128:256].unsqueeze(1), hs[:, :-1]), dim=1).reshape(B * 16, 128)
        _acc_grad(bp.lstm.weight_hh, dense(dxw, h_prev, a_trans=True))
        _acc_grad(bp.lstm.weight_ih, dense(dxw, tin.view(B * 16, 128), a_trans=True))
        dbg = _zeros((512 + 384,), dev)
        _lib.check(L.sb_colsum_f32(dxw.data_ptr(), B * 16, 512, 512, dbg.data_ptr(), st), 'sb_colsum_f32')
        _acc_grad(bp.lstm.bias_ih, dbg[:512])
        _acc_grad(bp.lstm.bias_hh, dbg[:512])
        dtin = dense(dxw, bp.lstm.weight_ih)  # [16B,128]
        _lib.check(L.sb_teacher_embed_bwd(dtin.data_ptr(), tints.data_ptr(), _ptr(keep1), 0.1, seed_ptr, seed_val, B,
                                          dproj.data_ptr(), 384, dw4.data_ptr(), db4.data_ptr(), st),
                   'sb_teacher_embed_bwd')
        _acc_grad(bp.dense4.weight, dw4)
        _acc_grad(bp.dense4.bias, db4)
        # ---- dense1 | dense2 | dense3 of the BitsPredictor
        lp.w123.split_grad(dense(dproj, h_flat, a_trans=True))
        _lib.check(L.sb_colsum_f32(dproj.data_ptr(), B, 384, 384, dbg[512:].data_ptr(), st), 'sb_colsum_f32')
        lp.b123.split_grad(dbg[512:])
        dh = dense(dproj, lp.w123.flat, out=dh_a)
        return dh, dx12, None, None, None, None, None


@torch.no_grad()
def latent_eval(h_flat, lp, rand, aux):
    """StochasticModel.forward, inference branch (:199-200): sampled bits + add_bits.  Returns (hm, x_mid)."""
    L, st = _lib.lib(), _lib.stream_ptr()
    B, dev = h_flat.shape[0], h_flat.device
    sm = lp.model.stochastic_model
    bp = sm.bits_predictor
    seed_ptr, seed_val = _seed(rand)
    proj = dense(h_flat, lp.w123.flat, b_trans=True, bias=lp.b123.flat)
    tb = bp._sampler_tables if (bp.cache_sampler_tables and bp._sampler_tables is not None) else None
    if tb is None:
        from . import ops
        tb = ops.lstm_sampler_tables(bp.lstm, bp.dense5, bp.dense4)
    noise = rand.sample_noise(B, 16) if hasattr(rand, 'sample_noise') else None
    bits = _sample_bits(proj[:, :128], proj[:, 128:256], proj[:, 256:384], 384, tb['bias'], tb['wihT'], tb['whhT'], tb['w5T'],
                        tb['E'], tb['b5'], noise, seed_ptr, B, L, st)
    force = getattr(rand, 'force', None)
    if force and 'sampled_bits' in force:  # test hook: decisions recorded from the oracle
        f = force['sampled_bits'].to(dev)
        rand.mismatch['sampled_bits'] = float((f != bits).float().mean())
        bits = f.contiguous()
    zz = dense(bits, lp.w23.flat, b_trans=True, bias=lp.b23.flat)
    hm = torch.empty(B, 3, 2, 768, device=dev, dtype=torch.bfloat16)
    x_mid = torch.empty(B, 768, device=dev)
    _lib.check(L.sb_latent_out_fwd(h_flat.data_ptr(), zz.data_ptr(), None, 0, None, 0.0, 0.0, seed_ptr, seed_val, B,
                                   hm.data_ptr(), x_mid.data_ptr(), None, st), 'sb_latent_out_fwd')
    if aux is not None:
        aux['bits'] = bits
    return hm, x_mid


# ---------------------------------------------------------------------------------------------------------------
class HeadsOut(torch.autograd.Function):
    """reward_pred = RewardEstimator(cat(x_mid, x_fin)); with targets also the reward CE and value MSE of
    simple/trainer.py:138-141 and every gradient of this block (computed by the forward launch: both losses enter the
    total with weight 1).  Returns (reward_pred, reward CE, value MSE)."""

    @staticmethod
    def forward(ctx, x_mid, x_fin, value_pred, rewards, values, model):
        L, st = _lib.lib(), _lib.stream_ptr()
        B, dev = x_mid.shape[0], x_mid.device
        re = model.reward_estimator
        r_in = torch.cat((x_mid, x_fin.unsqueeze(0).expand(B, -1)), dim=1)  # [B,864]
        hid = dense(r_in, re.dense1.weight, b_trans=True, bias=re.dense1.bias)       # [B,128]
        reward_pred = torch.empty(B, 3, device=dev)
        if rewards is None:
            _lib.check(L.sb_reward_value(hid.data_ptr(), re.dense2.weight.data_ptr(), re.dense2.bias.data_ptr(), None,
                                         None, None, B, reward_pred.data_ptr(), None, None, None, None, None, st),
                       'sb_reward_value')
            z = _zeros((2,), dev)
            return reward_pred, z[0], z[1]
        z = torch.zeros(2 + 3 * 128 + 3 + 128, device=dev)  # losses | dw2 | db2 | db1 (own buffer: the losses are returned)
        losses, dw2, db2 = z[:2], z[2:386].view(3, 128), z[386:389]
        dhid = torch.empty(B, 128, device=dev)
        dvalue = torch.empty(B, device=dev)
        vp = value_pred.detach().contiguous().float()
        _lib.check(L.sb_reward_value(hid.data_ptr(), re.dense2.weight.data_ptr(), re.dense2.bias.data_ptr(),
                                     rewards.contiguous().data_ptr(), vp.data_ptr(), values.contiguous().float().data_ptr(),
                                     B, reward_pred.data_ptr(), losses.data_ptr(), dhid.data_ptr(), dw2.data_ptr(),
                                     db2.data_ptr(), dvalue.data_ptr(), st), 'sb_reward_value')
        ctx.save_for_backward(r_in, dhid, dvalue, z)
        ctx.model = model
        ctx.mark_non_differentiable(reward_pred)
        return reward_pred, losses[0], losses[1]

    @staticmethod
    def backward(ctx, _, dl_reward, dl_value):
        r_in, dhid, dvalue, z = ctx.saved_tensors
        re = ctx.model.reward_estimator
        L, st = _lib.lib(), _lib.stream_ptr()
        B = r_in.shape[0]
        dw2, db2, db1 = z[2:386].view(3, 128), z[386:389], z[389:517]
        _acc_grad(re.dense2.weight, dw2)
        _acc_grad(re.dense2.bias, db2)
        _acc_grad(re.dense1.weight, dense(dhid, r_in, a_trans=True))
        _lib.check(L.sb_colsum_f32(dhid.data_ptr(), B, 128, 128, db1.data_ptr(), st), 'sb_colsum_f32')
        _acc_grad(re.dense1.bias, db1)
        dr_in = dense(dhid, re.dense1.weight)  # [B,864]
        dx_fin = _zeros((96,), r_in.device)
        _lib.check(L.sb_colsum_f32(dr_in[:, 768:].data_ptr(), B, 96, 864, dx_fin.data_ptr(), st), 'sb_colsum_f32')
        return dr_in[:, :768], dx_fin, dvalue, None, None, None
