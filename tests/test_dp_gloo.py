"""Data-parallel plumbing on CPU: world_size 2 over gloo (the N>1 path of bench.py uses the same code over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from simple_b200.parallel import FlatGradSync, broadcast_parameters, init_distributed
    r, w, _ = init_distributed('gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)  # different initial weights per rank ...
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    broadcast_parameters(model, src=0)  # ... made identical
    sync = FlatGradSync(model)
    params = list(model.parameters())
    for step in range(2):
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1 + i + 10 * step))
        if step == 1:
            params[0].grad = None  # like gate.* on the first step of a window: absent on every rank
        ran = []
        views = sync(model, None, lambda: ran.append(step))  # gradient-independent tail work of the caller
        assert ran == [step]  # ... is run exactly once per call (behind the last all-reduce on the GPU path)
        res = {i: views[p].clone() for i, p in enumerate(params) if p in views}
        out.put((rank, step, {i: v for i, v in res.items()}, [p.detach().clone() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_sync_averages_over_ranks():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world * 2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by = {(r, s): (g, w) for r, s, g, w in got}
    for step in range(2):
        g0, w0 = by[(0, step)]
        g1, w1 = by[(1, step)]
        assert all(torch.equal(a, b) for a, b in zip(w0, w1))  # identical replicas
        assert set(g0) == set(g1) == ({0, 1, 2, 3} if step == 0 else {1, 2, 3})
        for i in g0:
            want = (1 + i + 10 * step + 2 + i + 10 * step) / 2.0  # mean over the two ranks
            assert torch.equal(g0[i], g1[i]) and float(g0[i].flatten()[0]) == want


def _median_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from types import SimpleNamespace
    from simple_b200.parallel import init_distributed
    from simple_b200.trainer import Trainer
    init_distributed('gloo')
    g = torch.Generator().manual_seed(5)
    full = torch.randint(0, 256, (4, 12, 21, 16), generator=g, dtype=torch.uint8)  # even element count: lower median
    full[:2] //= 3  # the two shards have different distributions: a per-shard median would differ
    shard = full[rank * 2:(rank + 1) * 2]
    med = Trainer.global_median(SimpleNamespace(grad_sync=object()), shard)
    out.put((rank, float(med), float(torch.median(full.float() / 255)), float(torch.median(shard.float() / 255))))
    dist.barrier()
    dist.destroy_process_group()


def test_global_batch_median_by_histogram_all_reduce():
    """simple/trainer.py:64 takes torch.median over the WHOLE batch; sharded, that is a 256-bin histogram all-reduce."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_median_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, med, want, local in got:
        assert med == want, (rank, med, want)
    assert {local for *_, local in got} != {got[0][2]}  # (the shards' own medians are not the global one)


def test_gradient_arena_bucket_layout():
    """Host logic of the overlapped all-reduce: slots follow the backward completion order, a bucket closes after each
    boundary layer, and only the small shallow-encoder layers are left for the exposed last all-reduce."""
    from simple_b200.parallel import FlatGradSync
    sizes = {'down0': 3, 'down1': 12, 'down2': 47, 'down3': 94, 'down4': 94, 'mid0': 53, 'mid1': 53, 'up0': 94, 'up1': 94,
             'up2': 47, 'up3': 12, 'up4': 3, 'logits': 1, 'embed': 1, 'gate': 1, 's_embed': 1, 's_conv1': 8, 's_conv2': 42,
             'extra': 5}
    L = FlatGradSync.bucket_layout(sizes, FlatGradSync.BOUNDS_AFTER)
    assert L['names'][:6] == ['logits', 'up4', 'up3', 'up2', 'up1', 'up0'] and L['names'][-2:] == ['gate', 'extra']
    assert L['conv_total'] == sum(sizes.values())
    off = 0
    for n in L['names']:  # contiguous, non-overlapping slots
        assert L['slots'][n] == (off, sizes[n])
        off += sizes[n]
    assert [sorted(m) for m in L['members']] == [sorted(['logits', 'up4', 'up3', 'up2', 'up1', 'up0']),
                                                 sorted(['mid1', 'mid0', 's_conv2']), sorted(['s_conv1', 's_embed', 'down4']),
                                                 ['down3'], ['down2']]
    assert L['bounds'] == [251, 399, 502, 596, 643]
    left = L['conv_total'] - L['bounds'][-1]   # down1, down0, embed, gate (+ unknown layers): the exposed remainder
    assert left == 12 + 3 + 1 + 1 + 5 and left < 0.05 * L['conv_total']
