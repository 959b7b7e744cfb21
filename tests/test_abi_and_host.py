"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol include/simple_b200.h
declares (no compute calls without a GPU); host-side logic of the reference-facing surface."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from simple_b200 import _lib
    so = _lib.build()
    L = ctypes.CDLL(so)
    header = open(os.path.join(REPO, 'include', 'simple_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(sb_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 18, declared
    for name in sorted(declared):
        assert hasattr(L, name), f'{name} declared in include/simple_b200.h but not exported by {so}'
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    L.sb_version.restype = ctypes.c_char_p
    assert b'sm_100a' in L.sb_version()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import subprocess
    from simple_b200 import _lib
    so = _lib.build()
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):  # tcgen05.mma, TMA tensor loads, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert 'HMMA.16816' not in sass  # no legacy mma.sync path


def test_product_fails_loudly_without_cuda():
    from oracle.wm_oracle import default_config
    from simple_b200.adafactor import Adafactor
    from simple_b200.next_frame_predictor import NextFramePredictor
    from simple_b200.simulated_env import Discrete, make_simulated_env
    cfg = default_config(device='cpu')
    m = NextFramePredictor(cfg, 3)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 12, 105, 80), torch.zeros(1, 3))
    with pytest.raises(RuntimeError, match='CUDA'):
        make_simulated_env(cfg, m, Discrete(3))
    p = torch.nn.Parameter(torch.zeros(2, 2))
    p.grad = torch.ones(2, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        Adafactor([p]).step()
    with pytest.raises(ValueError):
        Adafactor([p], lr=0.1)                       # simple/adafactor.py:58-59
    with pytest.raises(ValueError):
        Adafactor([p], relative_step=False, warmup_init=True)  # simple/adafactor.py:60-61


def test_module_matches_reference_parameter_inventory():
    from oracle.wm_oracle import default_config
    from simple_b200.next_frame_predictor import NextFramePredictor
    torch.manual_seed(0)
    m = NextFramePredictor(default_config(device='cpu'), 3)
    sd = m.state_dict()
    assert len(sd) == 120 and sum(v.numel() for v in sd.values()) == 68340808
    golden = torch.load(os.path.join(REPO, 'tests', 'golden', 'wm_golden.pt'))
    assert list(sd.keys()) == list(golden['train'][0]['param_sample'].keys())
    assert sd['upscale_layers.0.weight'].shape == (768, 768, 4, 4) and sd['logits.weight'].shape == (768, 96, 1, 1)
    assert m.upscale_layers[0].output_padding == (0, 1) and m.upscale_layers[2].output_padding == (1, 0)
    assert m.upscale_layers[8].output_padding == (1, 0)


def test_static_plan_builds_on_cpu_and_wires_bias_gradient_side_channels():
    """The layer plan (layouts, conv specs, timing tables) is pure host logic: it must build without a GPU."""
    from oracle.wm_oracle import default_config
    from simple_b200.next_frame_predictor import NextFramePredictor
    m = NextFramePredictor(default_config(device='cpu'), 3)
    P = m._build_plan(torch.device('cpu'))
    assert len(P['down']) == 5 and len(P['up']) == 5 and len(P['prep_dec']) == 5
    assert P['logits'].dbias_src is m.logits_dbias
    fused = [P['embed'], P['s_embed'], P['mid'][1]] + P['down'] + P['up']
    assert all(sp.dbias_src == {} for sp in fused)
    assert len({id(sp.dbias_src) for sp in fused}) == len(fused)  # one private channel per conv
    # convs with several consumers of their output (or none that is a fused stage) keep the generic column-sum path
    assert all(sp.dbias_src is None for sp in (P['gate'], P['mid'][0], P['s_conv1'], P['s_conv2']))


def test_layouts_and_tiling_helpers():
    from simple_b200.gemm import choose_box
    from simple_b200.ops import s2d_in, s2d_out
    lay = s2d_in(105, 80, 96, 4, 2, 1)
    assert lay.shape(2) == (2, 53, 41, 384) and (lay.s, lay.pad) == (2, 1)
    lay4 = s2d_in(105, 80, 96, 8, 4, 2)
    assert lay4.shape(1) == (1, 27, 21, 1536)
    lo = s2d_out(6, 5, 13, 10, 768)
    assert lo.shape(3) == (3, 7, 6, 3072)
    x = torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 3, 5, 7)
    l2 = s2d_in(5, 7, 3, 4, 2, 1)
    full = torch.zeros(2, 3, l2.H2 * 2, l2.W2 * 2)
    full[:, :, 1:6, 1:8] = x
    phys = full.reshape(2, 3, l2.H2, 2, l2.W2, 2).permute(0, 2, 4, 3, 5, 1).reshape(l2.shape(2))
    assert torch.equal(l2.to_nchw(phys), x)
    for W, H, B in ((80, 105, 64), (2, 3, 16), (10, 13, 64), (41, 53, 7)):
        bw, bh, bb = choose_box(W, H, B)
        assert bw * bh * bb <= 128 and bw <= W and bh <= H
        bw, bh, bb = choose_box(W, H, B, rows=64, exact=True)
        assert bw * bh * bb == 64


def test_scheduled_sampling_epsilon_matches_reference_values():
    """SURVEY 9.11 (verified against simple/trainer.py:78-87): eps(0)=0.9999, eps(5000)=0.854, eps(10000)=0.545,
    eps(20000)=0.100, eps(>=22250)=0; epochs > 0 always feed back the model's own frame."""
    from types import SimpleNamespace
    from simple_b200.trainer import Trainer
    t = Trainer.__new__(Trainer)
    t.config = SimpleNamespace(scheduled_sampling_decay_steps=22250)
    for i, want in ((0, 0.9999), (5000, 0.854), (10000, 0.545), (20000, 0.100), (22250, 0.0), (40000, 0.0)):
        assert abs(t.epsilon_schedule(0, i) - want) < 1.5e-3, (i, t.epsilon_schedule(0, i))
    assert t.epsilon_schedule(3, 100) == 0


def test_device_replay_window_validity_and_agent_sharding():
    from simple_b200.simulated_env import shard_agents
    from simple_b200.trainer import DeviceReplay
    n, L = 30, 5
    buf = []
    for i in range(n):
        done = torch.tensor(1 if i == 12 else 0, dtype=torch.uint8)
        value = None if i >= 26 else torch.tensor(float(i))
        buf.append([torch.zeros(12, 4, 4, dtype=torch.uint8), torch.tensor([1, 0, 0], dtype=torch.uint8),
                    torch.tensor(1, dtype=torch.uint8), torch.zeros(3, 4, 4, dtype=torch.uint8), done, value])
    rp = DeviceReplay(buf, 'cpu', L)
    valid = set(rp.valid_starts.tolist())
    want = {s for s in range(n - L) if all(not (j == 12 or j >= 26) for j in range(s, s + L))}
    assert valid == want  # trainer.py:44-53: no terminal / value-less record among the L successors
    torch.manual_seed(0)
    assert set(rp.sample_starts(200).tolist()) <= valid
    assert [list(shard_agents(256, r, 8)) for r in (0, 7)] == [list(range(32)), list(range(224, 256))]
    assert sum(len(shard_agents(10, r, 4)) for r in range(4)) == 10


def test_separable_timing_table_equals_reference_signal():
    from oracle.wm_oracle import timing_signal_nd
    from simple_b200.next_frame_predictor import timing_signal_sep
    for C, H, W in ((96, 105, 80), (768, 6, 5)):
        ref = timing_signal_nd((C, H, W))[0]  # [C,H,W]
        sep = timing_signal_sep(C, H, W)
        th, tw = sep[:H], sep[H:]
        full = torch.cat((th[:, None, :].expand(H, W, C // 2), tw[None, :, :].expand(H, W, C // 2)), dim=2)
        assert torch.allclose(full.permute(2, 0, 1), ref, atol=1e-6)


def test_split_k_sizing_and_transpose_table_layout():
    """Host-side launch geometry: split-K factors target ONE wave of CTAs (148 SMs), and the descriptor records of the
    single-launch operand transpose match the 32-byte C struct of include/simple_b200.h."""
    import re
    import numpy as np
    from simple_b200 import ops
    # deep 3x2 feature map at batch 64: 384 rows = 3 M tiles, N = 768 = 3 N tiles -> 9 tiles, 16 splits = 144 CTAs
    assert ops.auto_splits(64, 3, 2, 768, 192) == 16          # 144 <= 148: no CTA gets a second work item
    assert ops.auto_splits(64, 105, 80, 96, 2) == 1            # thousands of tiles: never split
    assert 1 <= ops.auto_splits(16, 3, 2, 768, 192) <= 32      # rollout batch: capped
    assert ops.auto_splits(1, 3, 2, 768, 4) <= 4               # never more splits than K steps
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'simple_b200.h')).read()
    m = re.search(r'typedef struct sb_transpose_desc \{(.*?)\} sb_transpose_desc;', hdr, re.S)
    assert m, 'sb_transpose_desc missing from the header'
    fields = [f.strip() for f in m.group(1).split(';') if f.strip()]
    assert fields == ['const void* fwd', 'void* tr', 'int Opad, PHI, T, first_tile']
    dt = np.dtype([('fwd', np.uint64), ('tr', np.uint64), ('Opad', np.int32), ('PHI', np.int32), ('T', np.int32),
                   ('first_tile', np.int32)])
    assert dt.itemsize == 32


def test_host_side_launch_sizing_entry_points():
    """Pure host logic behind the C ABI (no GPU): K-split sizing of sb_dense, the SM-reserve switch."""
    from simple_b200 import _lib
    L = _lib.lib()
    # big outputs: one CTA per 64x64 tile, no split; small outputs: ~2 CTAs per SM, at least two 16-wide K tiles each
    assert L.sb_dense_splits(384, 4608, 64) == 1          # 6 x 72 tiles
    assert L.sb_dense_splits(64, 14976, 3) == 1           # 234 tiles
    s = L.sb_dense_splits(64, 384, 4608)                   # BitsPredictor dense1|2|3: 6 tiles
    assert s == 50 and 6 * s >= 148 * 2
    assert L.sb_dense_splits(64, 128, 60) == 2            # K = 60: two K tiles of 32 at most
    assert L.sb_dense_splits(1, 1, 1) == 1
    prev = L.sb_set_sm_reserve(8)
    try:
        assert L.sb_set_sm_reserve(1000) == 8             # returns the previous value; clamps
        assert L.sb_set_sm_reserve(-3) == 64
        assert L.sb_set_sm_reserve(0) == 0
    finally:
        L.sb_set_sm_reserve(prev)


def test_adafactor_grad_scale_rescales_second_moment_states():
    """Data parallel hands the optimiser SUMMED gradients (grad_scale = 1 / world): the states live in those units and
    are converted when the scale changes (a checkpoint moving between world sizes)."""
    from simple_b200.adafactor import Adafactor
    w = torch.nn.Parameter(torch.ones(4, 6))
    b = torch.nn.Parameter(torch.ones(6))
    opt = Adafactor([w, b])
    for p in (w, b):
        st = opt._init_state(p)
        for k in ('exp_avg_sq_row', 'exp_avg_sq_col', 'exp_avg_sq'):
            if k in st:
                st[k].fill_(2.0)
    assert opt.grad_scale == 1.0
    opt.grad_scale = 0.125                                   # gradients now arrive 8 x larger -> g^2 states 64 x
    assert float(opt.state[w]['exp_avg_sq_row'][0]) == 128.0 and float(opt.state[b]['exp_avg_sq'][0]) == 128.0
    opt.grad_scale = 0.5
    assert float(opt.state[w]['exp_avg_sq_col'][0]) == 8.0
    opt.grad_scale = 1.0
    assert float(opt.state[w]['exp_avg_sq_row'][0]) == 2.0
    with pytest.raises(ValueError):
        opt.grad_scale = 0.0
