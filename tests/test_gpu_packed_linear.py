"""GPU: packed-checkpoint path: MXQLinear routes to the GEMV / tcgen05 GEMM kernels
and equals x @ decode(pack(W))^T; save/load round trip is bit-exact."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _mlp(dev):
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(1024, 2048, bias=False), nn.Linear(2048, 512, bias=False)).to(dev).half()
    for lin in m:
        lin.weight.data.mul_(0.5)
    return m


@pytest.mark.parametrize("tokens", [(1,), (3,), (8,), (2, 5), (64,), (2, 300), (0,)])
def test_mxq_linear_matches_dequantized_matmul(cuda, tokens):
    from mxq_b200 import ops
    from mxq_b200.packed_linear import pack_linear
    lin = _mlp(cuda)[0]
    q = pack_linear(lin)
    x = torch.randn(*tokens, 1024, device=cuda).half()
    y = q(x)
    assert y.shape == (*tokens, 2048) and y.dtype == torch.float16
    if x.numel():
        ref = x.float().reshape(-1, 1024) @ ops.unpack(q.packed).T
        err = float((y.float().reshape(-1, 2048) - ref).abs().max() / ref.abs().max())
        assert err <= TOL
        # ... and approximates the unquantized layer (3.0 bits/weight)
        full = lin(x).detach().float().reshape(-1, 2048)
        assert float((ref - full).norm() / full.norm()) < 0.35


def test_convert_save_load_round_trip(cuda, tmp_path):
    from mxq_b200 import ops
    from mxq_b200.packed_linear import MXQLinear, convert_model, load_packed, save_packed
    m = _mlp(cuda)
    for lin in m:
        lin.mxq_packed = ops.pack(lin.weight.data)           # what nas_quant(args.pack=True) attaches
    x = torch.randn(16, 1024, device=cuda).half()
    conv = convert_model(m)
    assert all(isinstance(l, MXQLinear) for l in conv)
    y0 = conv(x)
    f = str(tmp_path / "packed.pt")
    save_packed(conv, f)
    fresh = load_packed(_mlp(cuda), f)
    assert all(isinstance(l, MXQLinear) for l in fresh)
    for a, b in zip(conv, fresh):
        for k in a.packed:
            assert torch.equal(a.packed[k], b.packed[k]), k
    assert torch.equal(fresh(x), y0)
    sd = fresh.state_dict()
    assert set(k.split(".")[-1] for k in sd) == set(a.packed)


def test_cpu_input_raises(cuda):
    from mxq_b200.packed_linear import MXQLinear
    q = MXQLinear(64, 64, device=cuda)
    with pytest.raises(RuntimeError, match="CUDA"):
        q(torch.zeros(1, 64))
