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


def test_importance_allocation_folded_into_the_packed_path(cuda):
    """the 4-bit group of every 64 columns chosen by the Wanda metric (prune.py:177) instead of by
    position, as a permutation of 16-column groups applied to the weights before packing, to x inside
    the decode GEMV's staging and before the prefill GEMM.  Packed tensors == the oracle's packer on the
    permuted weight; both kernels == the oracle on the permuted activations; and the important groups
    really end up with the finer grid."""
    import numpy as np
    from oracle import mxq_oracle as O
    from mxq_b200 import ops
    from mxq_b200.packed_linear import MXQLinear, pack_linear
    torch.manual_seed(0)
    OC, IC = 256, 1024
    W = (torch.randn(OC, IC, device=cuda) * 0.02).half()
    stat = torch.rand(IC, device=cuda) + 0.1
    hot = [3, 17, 40, 62]                          # one 16-column group made important in several blocks
    for g in hot:
        stat[16 * g:16 * g + 16] *= 400.0
    stat[5] = 0                                     # a dead column
    lin = torch.nn.Linear(IC, OC, bias=False).to(cuda).half()
    lin.weight.data = W
    m = pack_linear(lin, stat, importance=True)
    perm = m.group_perm.cpu().numpy()
    assert sorted(perm.tolist()) == list(range(IC // 16))
    for g in hot:
        assert perm[(g // 4) * 4 + 3] == g          # the hot group sits in its block's 4-bit slot
    colperm = (perm[:, None] * 16 + np.arange(16)).reshape(-1)
    Wp = W.cpu().numpy()[:, colperm]
    dead = (stat.cpu().numpy() == 0)[colperm]
    want = O.pack_mxq(Wp, dead)
    got = {k: v.cpu().numpy() for k, v in m.packed.items()}
    for k in want:
        a = got[k].view(np.uint16) if got[k].dtype == np.float16 else got[k]
        b = want[k].view(np.uint16) if want[k].dtype == np.float16 else want[k]
        assert np.array_equal(a, b), k
    # decode GEMV (permutation inside the activation staging) and prefill GEMM (gathered activations)
    for M in (1, 3, 300):
        x = np.random.default_rng(M).standard_normal((M, IC)).astype(np.float16)
        ref = O.gemm_mxq_f32(x[:, colperm], want)
        y = m(torch.from_numpy(x).to(cuda)).cpu().numpy().astype(np.float64)
        Wd = np.abs(O.decode_mxq(want).astype(np.float64))
        bound = 1e-3 * (np.abs(x[:, colperm].astype(np.float64)) @ Wd.T) + np.abs(ref) * 2.0 ** -11
        assert (np.abs(y - ref) <= bound).all(), M
    # the dequantized weight comes back in the original column order, and the hot groups are 4-bit accurate
    Wd = m.dequantize(torch.float32).cpu().numpy()
    assert np.array_equal(Wd[:, colperm], O.decode_mxq(want))
    pos = MXQLinear.from_packed(ops.pack(W, stat))          # positional recipe, same weights
    err_imp = np.abs(Wd - W.float().cpu().numpy()).reshape(OC, -1, 16).mean(axis=(0, 2))
    err_pos = np.abs(pos.dequantize(torch.float32).cpu().numpy() - W.float().cpu().numpy()).reshape(OC, -1, 16).mean(axis=(0, 2))
    for g in hot:
        if g % 4 != 3:
            assert err_imp[g] < 0.5 * err_pos[g]
    # save / load keeps the permutation
    import os
    import tempfile
    from mxq_b200.packed_linear import load_packed, save_packed
    holder = torch.nn.Sequential(m)
    with tempfile.TemporaryDirectory() as d:
        save_packed(holder, os.path.join(d, "p.pt"))
        fresh = torch.nn.Sequential(torch.nn.Linear(IC, OC, bias=False).to(cuda).half())
        load_packed(fresh, os.path.join(d, "p.pt"))
    assert torch.equal(fresh[0].group_perm, m.group_perm)
    x1 = torch.randn(1, IC, device=cuda).half()
    assert torch.equal(fresh[0](x1), m(x1))


def test_decode_chain_matches_module_forward(cuda):
    """q/k/v sharing one activation plus an o_proj that reads q's output, as ONE launch: same numbers as the
    modules called one by one (both kernels add exact integer group sums; the fp32 summation order differs)."""
    from mxq_b200.packed_linear import decode_chain, pack_linear
    torch.manual_seed(3)
    lins = [pack_linear(torch.nn.Linear(512, 256, bias=False).half().to(cuda)) for _ in range(3)]
    o = pack_linear(torch.nn.Linear(256, 512, bias=False).half().to(cuda))
    x = torch.randn(512, device=cuda).half()
    ys = [torch.zeros(256, dtype=torch.float16, device=cuda) for _ in range(3)]
    yo = torch.zeros(512, dtype=torch.float16, device=cuda)
    chain = decode_chain([(lins[0], x, ys[0], -1), (lins[1], x, ys[1], -1), (lins[2], x, ys[2], -1), (o, ys[0], yo, 0)])
    for _ in range(2):                                   # buffers are refilled between runs
        x.copy_(torch.randn(512, device=cuda).half())
        chain.run()
        torch.cuda.synchronize()
        for lin, y in zip(lins, ys):
            want = lin(x[None, :])[0].float()
            assert torch.allclose(y.float(), want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))
        want = o(ys[0][None, :])[0].float()
        assert torch.allclose(yo.float(), want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))
    with pytest.raises(TypeError):
        decode_chain([(torch.nn.Linear(4, 4), x, ys[0], -1)])
