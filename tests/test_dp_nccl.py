"""Data-parallel correctness ON HARDWARE: two NCCL ranks, each holding half of a global batch,
must end up with the weights a single rank computes on the whole batch -- same weights, same inputs, same injected
randomness (every draw is generated for the GLOBAL batch from one seeded generator and sliced per rank) -- and the two
replicas must stay BIT-IDENTICAL without any re-broadcast.  Needs 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu


class SlicedRand:
    """'Randomness as input' for a batch shard: draws every tensor for the global batch, hands out rows [lo, hi)."""

    def __init__(self, seed, b_global, lo, hi, device):
        self.g = torch.Generator().manual_seed(seed)
        self.bg, self.lo, self.hi, self.device = b_global, lo, hi, device
        self.seed, self.force, self.mismatch = 0, {}, {}

    def _draw(self, shape, fn):
        full = fn((self.bg,) + tuple(shape[1:]))
        return full[self.lo:self.hi].contiguous().to(self.device)

    def keep_mask(self, shape, p):
        return self._draw(shape, lambda s: (torch.rand(s, generator=self.g) >= p).float())

    def big_keep_mask(self, shape, p):
        return self.keep_mask(shape, p).permute(0, 2, 3, 1).contiguous().to(torch.uint8)

    def uniform(self, shape):
        return self._draw(shape, lambda s: torch.rand(s, generator=self.g))

    def exponential(self, shape):
        return self._draw(shape, lambda s: torch.empty(s).exponential_(1, generator=self.g))

    def sample_noise(self, B, n):
        return torch.stack([self.exponential((B, 256)) for _ in range(n)], dim=1).contiguous()

    def truncnorm(self, shape):
        return self._draw(shape, lambda s: (0.2 * torch.randn(s, generator=self.g)).clamp(-0.4, 0.4))


def _run(rank, world, b_global, steps, dev):
    """`steps` optimiser steps on rows [lo, hi) of the global batch; returns (weights, gradient norms)."""
    from oracle.wm_oracle import default_config
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.synthetic import synth_replay
    from simple_b200.trainer import Trainer
    per = b_global // world
    lo, hi = rank * per, (rank + 1) * per
    cfg = default_config(device=str(dev), batch_size=per, input_noise=0.0)
    torch.manual_seed(0)
    model = NextFramePredictor(cfg, 3)
    # well-conditioned variant (tests/test_strict_parity.py::_setup): on the reference init the global gradient norm --
    # and with it every update, through the clip coefficient -- swings by tens of percent between two executions of the
    # SAME step (fp32 summation order x bf16 storage x the 1000x gain of the 6-pixel InstanceNorms), single GPU included
    with torch.no_grad():
        for k in ('downscale_layers.6.bias', 'downscale_layers.8.bias', 'middle_network.middle_network.0.bias',
                  'middle_network.middle_network.2.bias', 'upscale_layers.0.bias'):
            dict(model.named_parameters())[k].add_(1.0)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    trainer = Trainer(model, cfg)
    if world > 1:
        from simple_b200.parallel import FlatGradSync
        trainer.grad_sync = FlatGradSync(model)
    d = synth_replay(b_global + steps + 2, 3, 11)
    frames = (d['obs'][lo:hi].float() / 255).to(dev)
    model.train()
    model.init_internal_states(per)
    norms = []
    for j in range(steps):
        idx = torch.arange(lo, hi) + j
        model.parity_rand = SlicedRand(100 + j, b_global, lo, hi, dev)
        out = trainer.train_step(frames, d['action'][idx].float().to(dev), d['new_obs'][idx].to(dev),
                                 d['reward'][idx].long().to(dev), d['value'][idx].to(dev), 0.0)
        norms.append(float(trainer.optimizer.grad_norm_sq))
        frames = torch.cat((frames[:, 3:], d['new_obs'][idx].to(dev).float() / 255), dim=1)  # teacher frames: same inputs
    model.parity_rand = None
    torch.cuda.synchronize()
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}, norms, init


def _worker(rank, world, port, b_global, steps, outdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from simple_b200.parallel import init_distributed
    init_distributed('nccl')
    dev = torch.device('cuda', rank)
    sd, norms, _ = _run(rank, world, b_global, steps, dev)
    torch.save({'sd': sd, 'norms': norms}, os.path.join(outdir, f'rank{rank}.pt'))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)  # (no communicator teardown underneath pending work; same exit path as bench.py)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_ranks_equal_one_rank_on_the_global_batch():
    import torch.multiprocessing as mp
    b_global, steps, world = 4, 3, 2
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context('spawn')
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, b_global, steps, d)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
            assert p.exitcode == 0
        r0, r1 = (torch.load(os.path.join(d, f'rank{r}.pt')) for r in range(world))
    # replicas: bit-identical after 3 steps, no broadcast in between
    for k in r0['sd']:
        assert torch.equal(r0['sd'][k], r1['sd'][k]), f'replicas diverged at {k}'
    assert r0['norms'] == r1['norms']
    one, norms1, init = _run(0, 1, b_global, steps, torch.device('cuda', 0))
    again, norms1b, _ = _run(0, 1, b_global, steps, torch.device('cuda', 0))
    noise_only = ('downscale_layers.6.bias', 'downscale_layers.8.bias', 'middle_network.middle_network.0.bias',
                  'middle_network.middle_network.2.bias', 'upscale_layers.0.bias')

    def update_diff(a, b):
        """|| update_a - update_b || / || update_a || over all parameters (global L2) after `steps` optimiser steps"""
        num = den = 0.0
        for k, w0 in init.items():
            if k in noise_only:
                continue
            num += float(((a[k] - b[k]) ** 2).sum())
            den += float(((a[k] - w0) ** 2).sum())
        return (num / den) ** 0.5

    floor = update_diff(one, again)      # two single-GPU runs of the same three steps (fp32 atomics x bf16 storage)
    dp = update_diff(r0['sd'], one)      # two ranks on the sharded batch vs one rank on the whole batch
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/dp_2rank_vs_1rank.txt', 'w') as f:
        f.write('gradient norms^2 per step: 2-rank %s | 1-rank %s | 1-rank again %s\n' % (r0['norms'], norms1, norms1b))
        f.write('update difference (global L2, relative): 2-rank vs 1-rank %.4e; 1-rank vs 1-rank (noise floor) %.4e\n'
                % (dp, floor))
        f.write('replicas bit-identical after %d steps without re-broadcast: yes\n' % steps)
    print(open('gpurun_out/dp_2rank_vs_1rank.txt').read())
    for a, b in zip(r0['norms'], norms1):
        assert abs(a - b) < 0.05 * abs(b), (r0['norms'], norms1)
    assert dp < max(0.03, 3 * floor), (dp, floor)
