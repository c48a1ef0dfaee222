"""GPU parity: calibration statistics + fused fasterquant vs the reference's golden vectors and the
oracle.  Bar: fake-quant fp16 weights and integer codes bit-exact; dead-column mask exact; running
statistics within 1e-5 relative (different fp32 summation order than torch.norm / X^T X)."""
import os

import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fqt(golden_dir):
    return np.load(os.path.join(golden_dir, "fasterquant.npz"))


def _linear(W16: np.ndarray, dev):
    N, K = W16.shape
    layer = torch.nn.Linear(K, N, bias=False)
    layer.weight.data = torch.from_numpy(W16.copy())
    return layer.to(dev)


def test_golden_mxqgpt(cuda, fqt):
    from mxq_b200 import MXQGPT, WrappedGPT
    for key in sorted({k.split("/")[0] for k in fqt.files}):
        W, X, Wq = fqt[key + "/W"], fqt[key + "/X"], fqt[key + "/Wq"]
        layer = _linear(W, cuda)
        gpt, wr = MXQGPT(layer), WrappedGPT(layer)
        assert (gpt.rows, gpt.columns, gpt.nsamples) == (W.shape[0], W.shape[1], 0)
        Xd = torch.from_numpy(X).to(cuda)
        for j in range(X.shape[0]):
            gpt.add_batch(Xd[j], None)
            wr.add_batch(Xd[j], None)
            ref = fqt[key + "/scaler_row"][j]
            got = wr.scaler_row.cpu().numpy()
            assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), key
        dH = gpt.diagH.cpu().numpy()
        assert np.array_equal(dH == 0, fqt[key + "/diagH"] == 0), "dead-column mask"
        assert np.abs(dH - fqt[key + "/diagH"]).max() <= 1e-5 * np.abs(dH).max()
        wm = wr.metric().cpu().numpy()
        assert np.abs(wm - fqt[key + "/wanda"]).max() <= 1e-5 * np.abs(wm).max()
        gpt.fasterquant(percdamp=0.01, blocksize=16)
        got = layer.weight.data.cpu().numpy()
        assert got.dtype == np.float16
        assert np.array_equal(got.view(np.uint16), Wq.view(np.uint16)), f"fasterquant {key}"
        gpt.free()


def test_golden_quantizer(cuda, golden_dir):
    from mxq_b200 import Quantizer
    Q = np.load(os.path.join(golden_dir, "quantizer.npz"))
    for key in sorted({k.split("/")[0] for k in Q.files}):
        bits = int(key[1])
        x = torch.from_numpy(Q[key + "/x"]).to(cuda)
        q = Quantizer()
        q.configure(bits=bits, perchannel=True, sym=False, qq_scale_bits=4)
        q.find_params(x, weight=True)
        assert np.array_equal(q.quantize_dequantize(x).cpu().numpy(), Q[key + "/y"]), key
        assert np.array_equal(q.quantize(x).cpu().numpy().astype(np.uint8), Q[key + "/codes"])
        assert np.array_equal(q.scale.reshape(-1).cpu().numpy(), Q[key + "/scale"])
        assert np.array_equal(q.zero.reshape(-1).cpu().numpy(), Q[key + "/zero"])


@pytest.mark.parametrize("shape", [(256, 1024), (48, 11008), (16, 64)])
def test_seeded_vs_oracle_with_codes(cuda, shape):
    from mxq_b200 import ops
    rng = np.random.default_rng(shape[1])
    W = (rng.standard_normal(shape) * 0.02).astype(np.float16)
    W[0, :16] = 0.25
    X = rng.standard_normal((64, shape[1])).astype(np.float16)
    X[:, [5, shape[1] - 1]] = 0
    X[:, 9] *= 20
    want, aux = O.fasterquant(W, O.dead_columns(X), return_aux=True)
    stat = ops.colsumsq(torch.from_numpy(X).to(cuda), add_scale=2.0 / 1)
    ref = O.colsumsq(X) * 2
    assert np.abs(stat.cpu().numpy() - ref).max() <= 1e-5 * ref.max()
    Wq, codes = ops.ptq_quant(torch.from_numpy(W).to(cuda), stat, return_codes=True)
    assert np.array_equal(codes.cpu().numpy(), aux["codes"]), "integer codes"
    assert np.array_equal(Wq.cpu().numpy().view(np.uint16), want.view(np.uint16))
    # no statistics -> nothing is dead
    want2 = O.fasterquant(W, None)
    got2 = ops.ptq_quant(torch.from_numpy(W).to(cuda), None)
    assert np.array_equal(got2.cpu().numpy().view(np.uint16), want2.view(np.uint16))


@pytest.mark.parametrize("shape", [(4096, 4096), (4096, 11008), (11008, 4096)])
def test_full_size_properties(cuda, shape):
    """Llama-2-7B shapes: oracle on sampled 16-row tiles, tile independence, idempotent dead columns."""
    from mxq_b200 import ops
    torch.manual_seed(shape[1])
    W = (torch.randn(*shape, device=cuda) * 0.02).half()
    stat = torch.ones(shape[1], device=cuda)
    stat[7] = 0
    Wq = ops.ptq_quant(W, stat)
    # a dead column is zeroed BEFORE quantization (mxqgpt.py:403); with the float zero-point its
    # dequantized value is scale*(round(zero)-zero), small but not exactly 0 -- same as the reference
    assert float(Wq[:, 7].float().abs().max()) < float(W[:, 7].float().abs().max())
    tiles = np.random.default_rng(0).choice(shape[0] // 16, 4, replace=False)
    dead = np.zeros(shape[1], bool)
    dead[7] = True
    for t in tiles:
        sl = slice(16 * t, 16 * t + 16)
        want = O.fasterquant(W[sl].cpu().numpy(), dead)
        assert np.array_equal(Wq[sl].cpu().numpy().view(np.uint16), want.view(np.uint16))
    # 16-row tiles are independent: quantizing a slab alone reproduces the same rows
    assert torch.equal(ops.ptq_quant(W[160:320].contiguous(), stat), Wq[160:320])
    # at most 4 distinct values per (row, 2-bit group) and 16 per row pool
    g = Wq[:64].view(64, -1, 4, 16)[:, :, :3].float().cpu().numpy()
    assert max(len(np.unique(r)) for r in g.reshape(-1, 16)[:2000]) <= 4
    pool = Wq[:64].view(64, -1, 4, 16)[:, :, 3].reshape(64, -1).float().cpu().numpy()
    assert max(len(np.unique(r)) for r in pool) <= 16


def test_colsumsq_shapes_and_dtypes(cuda):
    from mxq_b200 import ops
    rng = np.random.default_rng(0)
    for dt, tol in ((torch.float16, 1e-5), (torch.bfloat16, 1e-5), (torch.float32, 1e-5)):
        for shape in ((3000, 4096), (777, 11008), (1, 64), (130, 8)):
            X = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(dt)
            got = ops.colsumsq(X.to(cuda)).cpu().numpy()
            ref = O.colsumsq(X.float().numpy())
            assert np.abs(got - ref).max() <= tol * max(ref.max(), 1e-30), (dt, shape)
    # running update: out = prev*out + add*sumsq
    X = torch.randn(100, 256, device=cuda).half()
    a = ops.colsumsq(X)
    b = a.clone()
    ops.colsumsq(X, out=b, prev_scale=0.5, add_scale=0.25)
    assert torch.allclose(b, 0.75 * a, rtol=1e-6)


@pytest.mark.parametrize("shape", [(64, 256), (4096, 4096), (1000, 11008), (16, 64)])
def test_allocate_group_bits_bit_exact(cuda, shape):
    """f-3: importance-driven 2/4-bit allocation mask and fp64 importances equal the oracle exactly
    (integer column sums + fixed-order fp64), and the mask drives the PTQ / fake-quant kernels."""
    from mxq_b200 import ops
    g = torch.Generator().manual_seed(shape[0] + shape[1])
    W = (torch.randn(*shape, generator=g) * 0.02).half()
    W[:, 16:32] *= 6
    sr = torch.rand(shape[1], generator=g) * 4
    sr[7] = 0
    for stat in (None, sr):
        want_gb, want_imp = O.allocate_group_bits(W.numpy(), None if stat is None else stat.numpy())
        gb, imp = ops.allocate_group_bits(W.to(cuda), None if stat is None else stat.to(cuda), return_importance=True)
        assert np.array_equal(imp.cpu().numpy(), want_imp), "importances (fp64, bit-exact)"
        assert np.array_equal(gb.cpu().numpy(), want_gb), "allocation mask"
    if shape[0] % 16 == 0:
        out = ops.fakequant_fwd(W.to(cuda).float(), group_bits=gb)
        want = O.fakequant_fwd(W.float().numpy(), "fp32", 2, group_bits=want_gb)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_pipelined_layers_equal_serial(cuda):
    """LlamaLayerPTQ.run_pipelined (statistics of layer i+1 on the current stream, capped at 5 CTAs per
    SM, overlapping quantize+pack of layer i on a side stream) == run() layer by layer, bit for bit,
    including the double-buffered statistics (dead column differs per layer)."""
    from mxq_b200 import prune
    hidden, inter, tokens, nl = 256, 512, 1024, 5
    ptq = prune.LlamaLayerPTQ(hidden, inter, cuda, tokens)
    lin = prune.llama_linears(hidden, inter)
    g = torch.Generator(device=cuda).manual_seed(0)
    layers = []
    for l in range(nl):
        calib = {}
        for key, d in (("attn_in", hidden), ("o_in", hidden), ("mlp_in", hidden), ("down_in", inter)):
            X = torch.randn((tokens, d), generator=g, device=cuda, dtype=torch.float16)
            X[:, 3 + l] = 0
            calib[key] = X
        weights = {name: (torch.randn((oc, ic), generator=g, device=cuda) * 0.02).half() for name, (oc, ic, _) in lin.items()}
        layers.append((weights, calib))

    def collect(runner):
        got = []
        def sink(name, Wq, packed):
            got.append((name, Wq.clone(), {k: v.clone() for k, v in packed.items()}))
        runner(sink)
        torch.cuda.synchronize()
        return got

    serial = collect(lambda sink: [ptq.run(w, c, 8, sink=sink) for w, c in layers])
    piped = collect(lambda sink: ptq.run_pipelined(layers, 8, sink=sink))
    assert len(serial) == len(piped) == nl * 7
    for (n0, w0, p0), (n1, w1, p1) in zip(serial, piped):
        assert n0 == n1 and torch.equal(w0, w1)
        for k in p0:
            assert torch.equal(p0[k], p1[k]), (n0, k)
    # the occupancy-capped statistics kernel gives the same sums as the default one
    X = layers[0][1]["down_in"]
    from mxq_b200 import ops
    ref = ops.colsumsq(X)
    for ctas in (0, 1, 3, 5, 7):      # 0 = the shared-memory ring kernel
        out = torch.empty_like(ref)
        ws = torch.empty(ops.L.lib().mxq_colsumsq_workspace_bytes(X.shape[0], X.shape[1]), dtype=torch.uint8, device=cuda)
        rc = ops.L.lib().mxq_colsumsq_ex(X.data_ptr(), X.shape[0], X.shape[1], ops.L.MXQ_F16, out.data_ptr(), 0.0, 1.0, 0,
                                         ctas, ws.data_ptr(), ws.numel(), ops.L.stream())
        assert rc == 0
        assert torch.allclose(out, ref, rtol=1e-5, atol=0)


def test_colsumsq_ring_kernel(cuda):
    """mxq_colsumsq_ex(ctas_per_sm=0): TMA bulk ring, one CTA per SM -- same sums as the oracle for every
    chunk count (1..8 chunks per thread), ragged last stages, rows of one stage or several."""
    from mxq_b200 import ops
    rng = np.random.default_rng(3)
    L = ops.L
    for dt, code, shape in ((torch.float16, L.MXQ_F16, (3000, 4096)), (torch.bfloat16, L.MXQ_BF16, (777, 11008)),
                            (torch.float32, L.MXQ_F32, (1000, 1024)), (torch.float32, L.MXQ_F32, (2500, 8192)),
                            (torch.float16, L.MXQ_F16, (149, 64)), (torch.float16, L.MXQ_F16, (4099, 16384)),
                            (torch.float16, L.MXQ_F16, (32768, 4096))):
        X = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(dt)
        X[:, 7] = 0
        Xd = X.to(cuda)
        out = torch.full((shape[1],), -1.0, dtype=torch.float32, device=cuda)
        ws = torch.empty(L.lib().mxq_colsumsq_workspace_bytes(shape[0], shape[1]), dtype=torch.uint8, device=cuda)
        rc = L.lib().mxq_colsumsq_ex(Xd.data_ptr(), shape[0], shape[1], code, out.data_ptr(), 0.0, 1.0, 0, 0,
                                     ws.data_ptr(), ws.numel(), L.stream())
        assert rc == 0
        ref = O.colsumsq(X.float().numpy())
        got = out.cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-5 * ref.max(), (dt, shape)
        assert got[7] == 0.0 and np.array_equal(got == 0, O.dead_columns(X.float().numpy()))


@pytest.mark.parametrize("bs", [128, 48, 32])
def test_fasterquant_default_and_other_blocksizes(cuda, golden_dir, bs):
    """MXQGPT.fasterquant with the signature's default blocksize=128 (mxqgpt.py:388: one 48-wide 2-bit
    group per 64-column block, :413-415) and 48 / 32, bit-exact vs the unmodified reference's output."""
    import os
    from mxq_b200 import MXQGPT
    d = np.load(os.path.join(golden_dir, "fasterquant_blocksize.npz"))
    W, X, Wq = d[f"bs{bs}/W"], d[f"bs{bs}/X"], d[f"bs{bs}/Wq"]
    layer = torch.nn.Linear(W.shape[1], W.shape[0], bias=False).to(cuda).half()
    layer.weight.data = torch.from_numpy(W).to(cuda)
    gpt = MXQGPT(layer)
    for j in range(X.shape[0]):
        gpt.add_batch(torch.from_numpy(X[j]).to(cuda), None)
    if bs == 128:
        gpt.fasterquant()                                  # the default
    else:
        gpt.fasterquant(blocksize=bs)
    got = layer.weight.data.cpu().numpy()
    assert np.array_equal(got.view(np.uint16), Wq.view(np.uint16))
    with pytest.raises(NotImplementedError):
        MXQGPT(layer).fasterquant(blocksize=8)


def test_fasterquant_without_samples_zeroes_the_weight(cuda):
    """nsamples == 0: H == 0, every column is dead, W[:, dead] = 0 (mxqgpt.py:399-403)."""
    from mxq_b200 import MXQGPT
    torch.manual_seed(3)
    layer = torch.nn.Linear(256, 64, bias=False).to(cuda).half()
    W0 = layer.weight.data.cpu().numpy().copy()
    gpt = MXQGPT(layer)
    gpt.fasterquant(blocksize=16)
    want = O.fasterquant(W0, np.ones(256, bool))
    assert np.array_equal(layer.weight.data.cpu().numpy().view(np.uint16), want.view(np.uint16))
