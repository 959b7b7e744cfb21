"""sb_dense (skinny fp32 GEMMs of the latent / heads block) against torch fp32 matmul in float64 accumulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (M, N, K, a_trans, b_trans, bias, pitch padding): the shapes of the training step plus ragged / unaligned ones
CASES = [
    (64, 384, 4608, False, True, True, 0),      # BitsPredictor dense1|2|3: split-K 48 ways
    (64, 14976, 3, False, True, True, 0),       # ActionInjector bank, K = n_action
    (64, 14976, 9, False, True, True, 0),       # ... n_action = 9
    (14976, 3, 64, True, False, False, 0),      # its weight gradient, 3-wide output rows
    (384, 4608, 64, True, False, False, 0),     # dense1|2|3 weight gradient (store mode)
    (64, 4608, 384, False, False, False, 0),    # ... input gradient
    (1024, 512, 128, False, False, True, 0),    # LSTM input projection of all 16 steps
    (512, 128, 1024, True, False, False, 0),    # LSTM weight gradient, K = 16 B
    (64, 128, 60, False, True, True, 0),        # StochasticModel.dense1 on [B,60]
    (64, 128, 864, False, True, True, 0),       # reward MLP
    (64, 128, 128, False, False, False, 256),   # views into a wider buffer (proj[:, :128])
    (5, 7, 13, False, True, True, 3),           # ragged, unaligned pitches
    (130, 70, 33, True, True, True, 1),
    (1, 1, 1, False, False, False, 0),
]


@pytest.mark.parametrize('M,N,K,at,bt,use_bias,pad', CASES)
def test_dense_matches_float64(M, N, K, at, bt, use_bias, pad):
    from simple_b200 import ops
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)

    def mk(r, c):
        full = torch.randn(r, c + pad, generator=g).to(dev)
        return full[:, :c] if pad else full
    a = mk(K, M) if at else mk(M, K)
    b = mk(N, K) if bt else mk(K, N)
    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
    ref = (a.double().t() if at else a.double()) @ (b.double().t() if bt else b.double())
    if bias is not None:
        ref = ref + bias.double()
    out = ops.dense(a, b, a_trans=at, b_trans=bt, bias=bias)
    assert out.shape == (M, N)
    scale = float(ref.abs().max()) + 1e-12
    assert float((out.double() - ref).abs().max()) / scale < 2e-6
    # accumulate into an addend held in a wider buffer
    base = torch.randn(M, N + 4, generator=g).to(dev)
    acc = base.clone()
    ops.dense(a, b, a_trans=at, b_trans=bt, bias=bias, out=acc[:, :N])
    assert float((acc[:, :N].double() - (ref + base[:, :N].double())).abs().max()) / (scale + 4) < 2e-6
    assert torch.equal(acc[:, N:], base[:, N:])  # nothing written outside the [M,N] window


def test_dense_rejects_bad_operands():
    from simple_b200 import ops
    dev = torch.device('cuda:0')
    a, b = torch.randn(4, 5, device=dev), torch.randn(6, 7, device=dev)
    with pytest.raises(ValueError):
        ops.dense(a, b)
    with pytest.raises(ValueError):
        ops.dense(a.half(), torch.randn(5, 7, device=dev))
    with pytest.raises(ValueError):
        ops.dense(a, torch.randn(5, 7, device=dev), bias=torch.randn(6, device=dev))
