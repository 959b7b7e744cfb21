"""Diagnostic: run ONE training step four times from the same restored state (eager, eager, graph, graph) and report
what differs (losses, latent bits, gradient norm, weights).  python tools/diag_determinism.py"""
import os
import sys

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)
sys.path.insert(0, os.path.join(_REPO, 'tests'))
import test_graph_equivalence as T  # noqa: E402
from test_graph_equivalence import DEV, _Snapshot, _cfg, _replay  # noqa: E402


def main():
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.trainer import StepRunner, Trainer
    B = 4
    cfg = _cfg(batch_size=B, rollout_length=8)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3)
    if os.environ.get('SB_CONDITIONED', '1') != '0':
        with torch.no_grad():
            for k in ('downscale_layers.6.bias', 'downscale_layers.8.bias', 'middle_network.middle_network.0.bias',
                      'middle_network.middle_network.2.bias', 'upscale_layers.0.bias'):
                dict(model.named_parameters())[k].add_(1.0)
    model = model.to(DEV)
    trainer = Trainer(model, cfg)
    runner = StepRunner(trainer, _replay(B + 24), B, use_graph=True, resident=True)
    runner.begin(torch.arange(B, device=DEV), 0.3)
    for _ in range(5):
        runner.step()
    torch.cuda.synchronize()
    snap = _Snapshot(trainer, runner)
    runs = {}
    for name in ('eager1', 'eager2', 'graph1', 'graph2'):
        snap.restore()
        if name.startswith('graph'):
            runner.step()
        else:
            runner._body()
            runner.j += 1
        torch.cuda.synchronize()
        aux = model.last_aux
        runs[name] = dict(losses=runner.losses.clone().cpu(), gn=float(trainer.optimizer.grad_norm_sq),
                          bits=aux['bits'].clone().cpu() if 'bits' in aux else None,
                          pre=aux['posterior_pre'].clone().cpu() if 'posterior_pre' in aux else None,
                          params={k: p.detach().clone().cpu() for k, p in model.named_parameters()},
                          frames=runner.frames.clone().cpu(), rng=torch.cuda.get_rng_state(torch.device(DEV))[-16:].tolist())
        print(name, 'losses', runs[name]['losses'].tolist(), 'gn', runs[name]['gn'], 'rng tail', runs[name]['rng'])
    for a, b in (('eager1', 'eager2'), ('graph1', 'graph2'), ('eager1', 'graph1')):
        ra, rb = runs[a], runs[b]
        rows = sorted(((float((ra['params'][k] - rb['params'][k]).abs().max() / (rb['params'][k].abs().max() + 1e-12)),
                        float((ra['params'][k] - rb['params'][k]).norm() / (ra['params'][k] - snap_params[k]).norm().clamp(min=1e-30)), k)
                       for k in ra['params']), reverse=True)
        print(a, 'vs', b, 'loss diff', (ra['losses'] - rb['losses']).abs().max().item(),
              'bits equal', None if ra['bits'] is None else bool(torch.equal(ra['bits'], rb['bits'])),
              'pre maxdiff', None if ra['pre'] is None else float((ra['pre'] - rb['pre']).abs().max()),
              'frames diff frac', float((ra['frames'] != rb['frames']).float().mean()))
        for r in rows[:6]:
            print('    max|dw|/max|w| %.3e   |dw_a - dw_b| / |dw_a| %.3e   %s' % r)


if __name__ == '__main__':
    # weights before the step, to express differences relative to the size of the UPDATE
    snap_params = {}
    _orig = T._Snapshot.__init__

    def _init(self, trainer, runner):
        _orig(self, trainer, runner)
        snap_params.update({k: p.detach().clone().cpu() for k, p in trainer.model.named_parameters()})
    T._Snapshot.__init__ = _init
    main()
