import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_replay
from oracle.wm_oracle import default_config
from simple_b200.next_frame_predictor import NextFramePredictor
from simple_b200.ops import PackedWeight
from simple_b200.trainer import Trainer
DEV = 'cuda:0'
B = 2
cfg = default_config(device=DEV, batch_size=B)
d = {k: v.to(DEV) for k, v in synth_replay(B + 4, 3, 3).items()}
res = []
for packed in (False, True, False):
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3).to(DEV)
    trainer = Trainer(model, cfg)
    trainer.packed_grads = packed
    model.init_internal_states(B)
    frames = d['obs'][:B].float() / 255
    losses = []
    snaps = []
    for j in range(2):
        idx = torch.arange(B, device=DEV) + j
        out = trainer.train_step(frames, d['action'][idx].float(), d['new_obs'][idx], d['reward'][idx].long(), d['value'][idx], 0)
        frames = torch.cat((frames[:, 3:], out['argmax'].float() / 255), dim=1)
        losses.append(float(out['loss']))
        snaps.append({k: v.detach().clone() for k, v in model.named_parameters()})
    res.append((model, losses, snaps))
    print('packed', packed, 'losses', losses, 'gnorm', float(trainer.optimizer.grad_norm_sq))
for j in range(2):
    print('--- after step', j)
    for k in res[0][2][j]:
        a, b, c = res[0][2][j][k], res[1][2][j][k], res[2][2][j][k]
        e1 = float((a - b).abs().max() / (a.abs().max() + 1e-12))
        e2 = float((a - c).abs().max() / (a.abs().max() + 1e-12))
        if e1 > 1e-5 or e2 > 1e-5:
            print(k, 'ref-vs-packed %.3e' % e1, 'ref-vs-ref %.3e' % e2)
m1 = res[1][0]
mods = m1._weights()
for k, pw in m1._packed.items():
    ref = PackedWeight(mods[k].weight, pw.s, Ipad=pw.Ipad, Opad=pw.Opad, iperm=pw.iperm.tolist() if pw.iperm is not None else None,
                       operm=pw.operm.tolist() if pw.operm is not None else None)
    print(k, 'fwd eq', torch.equal(ref.fwd, pw.fwd), 'tr eq', torch.equal(ref.tr, pw.tr))
