"""Fused clip_grad_norm_ + Adafactor kernels vs the CPU oracle restatement of simple/adafactor.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_fused_adafactor_matches_oracle():
    from oracle import wm_oracle as O
    from simple_b200.adafactor import Adafactor
    torch.manual_seed(0)
    shapes = [(128, 76, 3, 3), (192, 96, 4, 4), (96, 76, 1, 1), (128, 32, 8, 8), (4, 128, 1, 1), (768, 3), (128, 4608),
              (60, 640), (3, 128), (1, 4608), (768,), (96,), (1,)]
    params = [torch.nn.Parameter((torch.randn(s) * 0.05).to(DEV)) for s in shapes]
    skip = 2  # this parameter has no gradient on the first step (like gate.* on rollout step 0)
    opt = Adafactor(params)
    cpu_p = [p.detach().cpu().clone() for p in params]
    cpu_state = [{} for _ in params]
    for step in range(3):
        grads = [torch.randn(s) * (10.0 if step == 0 else 0.01) for s in shapes]  # step 0 clips, later steps do not
        present = [i for i in range(len(shapes)) if not (step == 0 and i == skip)]
        for i, p in enumerate(params):
            p.grad = grads[i].to(DEV) if i in present else None
        opt.step(max_grad_norm=1.0)
        clipped, total = O.clip_grad_norm([grads[i] for i in present], 1.0)
        for g, i in zip(clipped, present):
            cpu_p[i] = O.adafactor_step(cpu_p[i], g, cpu_state[i])
        torch.cuda.synchronize()
        assert abs(float(opt.grad_norm_sq.sqrt()) - float(total)) < 1e-3 * float(total)
        for i, p in enumerate(params):
            ref = cpu_p[i]
            d = float((p.detach().cpu() - ref).abs().max() / (ref.abs().max() + 1e-12))
            assert d < 2e-5, (step, shapes[i], d)
            st = opt.state[p]
            if i in present or step > 0:
                assert st['step'] == cpu_state[i]['step']
                if len(shapes[i]) >= 2:
                    a, b = st['exp_avg_sq_row'].cpu(), cpu_state[i]['exp_avg_sq_row']
                    assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-12)
                    a, b = st['exp_avg_sq_col'].cpu(), cpu_state[i]['exp_avg_sq_col']
                    assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-12)
                else:
                    assert torch.allclose(st['exp_avg_sq'].cpu(), cpu_state[i]['exp_avg_sq'], rtol=1e-4, atol=1e-12)
                assert abs(float(st['RMS']) - float(cpu_state[i]['RMS'])) < 1e-5


def test_adafactor_rejects_non_default_and_cpu():
    from simple_b200.adafactor import Adafactor
    p = torch.nn.Parameter(torch.zeros(4, 4))
    with pytest.raises(ValueError):
        Adafactor([p], lr=1e-3)
    with pytest.raises(ValueError):
        Adafactor([p], beta1=0.9)
    opt = Adafactor([p])
    p.grad = torch.ones(4, 4)
    with pytest.raises(RuntimeError):
        opt.step()
