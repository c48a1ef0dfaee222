"""GPU parity: SymQuantizer / AsymQuantizer (activation + KV-cache fake quantizers)
vs the reference's golden vectors and the oracle.  Bar: bit-exact for fp32, bf16 and fp16."""
import os

import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O
from tests.gpu_util import TD, bits_equal, to_dev, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aq(golden_dir):
    return np.load(os.path.join(golden_dir, "actquant.npz"))


def test_golden_forward_backward(cuda, aq):
    from mxq_b200 import AsymQuantizer, SymQuantizer
    clip = torch.tensor([-2.0, 2.0])
    n = 0
    for k in sorted(k for k in aq.files if k.endswith("/y")):
        mode, dtype, case, b, _ = k.split("/")
        fn = SymQuantizer if mode == "sym" else AsymQuantizer
        x = to_dev(aq[f"{mode}/{dtype}/{case}/x"], dtype, cuda).requires_grad_(True)
        y = fn.apply(x, clip, int(b[1:]), case.startswith("layerwise"))
        assert y.dtype == x.dtype and y.shape == x.shape
        assert bits_equal(to_np(y), aq[k]), f"forward {k}"
        if k[:-2] + "/gi" in aq.files:
            y.backward(to_dev(aq[k[:-2] + "/go"], dtype, cuda))
            assert bits_equal(to_np(x.grad), aq[k[:-2] + "/gi"]), f"backward {k}"
        n += 1
    assert n >= 45


# every kernel family: sub-warp (2-D groups), one CTA per segment with 1..12 chunks per thread
# (3-D activations up to 11008 channels), and the two-kernel path (4-D per head, layerwise)
@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("mode,shape,layerwise", [
    ("sym", (333, 1024), False), ("asym", (257, 256), False),
    ("sym", (2, 256, 4096), False), ("asym", (1, 300, 11008), False), ("sym", (3, 100, 2048), False),
    ("asym", (2, 200, 128), False), ("sym", (1, 4200, 512), False),
    ("asym", (2, 8, 128, 128), False), ("sym", (1, 40, 64, 128), False),
    ("sym", (64, 4096), True), ("asym", (7, 33, 64), True),
])
def test_seeded_vs_oracle(cuda, dtype, mode, shape, layerwise):
    from mxq_b200 import AsymQuantizer, SymQuantizer
    g = torch.Generator().manual_seed(sum(shape) + len(dtype))
    x = (torch.randn(*shape, generator=g) * 1.5).to(TD[dtype])
    if mode == "sym" and len(shape) == 3 and shape[1] > (shape[2] // 128) * 128:
        pass    # tokens beyond (C // 128) * 128 exercise the zero-statistic quirk
    bits = 8 if mode == "sym" else 4
    fn_o = O.sym_quant if mode == "sym" else O.asym_quant
    want = fn_o(x.float().numpy(), dtype, bits, layerwise)
    fn = SymQuantizer if mode == "sym" else AsymQuantizer
    got = fn.apply(x.to(cuda), torch.tensor([-2.0, 2.0]), bits, layerwise)
    assert bits_equal(to_np(got), want)


def test_idempotent_and_range(cuda):
    """size-independent properties at a QAT-sized activation: [2, 2048, 4096] bf16.
    A symmetric 8-bit quantizer's output takes at most 255 distinct values per token and
    stays within the token's |x|max."""
    from mxq_b200 import SymQuantizer
    g = torch.Generator(device=cuda).manual_seed(0)
    x = torch.randn(2, 2048, 4096, generator=g, device=cuda, dtype=torch.bfloat16)
    clip = torch.tensor([-2.0, 2.0])
    y = SymQuantizer.apply(x, clip, 8, False)
    m = x.abs().amax(dim=-1, keepdim=True).float()
    assert (y.float().abs() <= m * 1.01).all()
    row = y[0, 5].float()
    assert row.unique().numel() <= 255
    # a slice of the big tensor equals the oracle on that slice (tokens are independent)
    want = O.sym_quant(x[1:2, 100:104].float().cpu().numpy(), "bf16", 8, False)
    assert bits_equal(to_np(y[1:2, 100:104]), want)


def test_quantize_linear_act_quant(cuda):
    """QuantizeLinear(w_bits=2, a_bits=8): weight fake-quant + activation fake-quant + linear
    (utils_quant.py:635-639,717-723), forward and backward through both STEs."""
    from mxq_b200 import QuantizeLinear
    torch.manual_seed(0)
    lin = QuantizeLinear(256, 128, w_bits=2, a_bits=8, symmetric=True).to(cuda)
    x = torch.randn(2, 16, 256, device=cuda, requires_grad=True)
    out = lin(x)
    wq = O.fakequant_fwd(lin.weight.detach().cpu().numpy(), "fp32", 2)
    xq = O.sym_quant(x.detach().cpu().numpy(), "fp32", 8, False)
    want = xq.reshape(-1, 256) @ wq.T
    assert np.allclose(to_np(out).reshape(-1, 128), want, rtol=1e-4, atol=1e-4)
    out.sum().backward()
    assert x.grad is not None and lin.weight.grad is not None
    assert torch.isfinite(x.grad).all() and torch.isfinite(lin.weight.grad).all()


def test_argument_errors(cuda):
    from mxq_b200 import ops
    x = torch.randn(4, 100, device=cuda)
    with pytest.raises(NotImplementedError):
        ops.segquant_plan(tuple(x.shape), "sym", False)      # 100 % 128 != 0
    with pytest.raises(RuntimeError):
        ops.segquant_fwd(torch.randn(4, 6, device=cuda), "asym", 4, 4, 6)   # 24 B segments
    with pytest.raises(RuntimeError):
        ops.segquant_fwd(torch.randn(4, 8), "asym", 4, 4, 8)               # CPU tensor
    e = ops.segquant_fwd(torch.empty(0, 8, device=cuda), "asym", 4, 0, 8)
    assert e.numel() == 0


@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("mode", ["sym", "asym"])
def test_non_finite_inputs_follow_ieee(cuda, dtype, mode):
    """inf / nan / huge / subnormal-tiny values take the guarded IEEE path and still equal the
    op-by-op oracle (NaNs in the same places): 2-D groups and 3-D per-token segments."""
    from mxq_b200 import AsymQuantizer, SymQuantizer
    g = torch.Generator().manual_seed(11)
    for shape in ((160, 256), (1, 160, 512)):
        x = torch.randn(*shape, generator=g)
        flat = x.view(-1, shape[-1])
        flat[1, 3] = float("inf")
        flat[2, 5] = float("-inf")
        flat[3, 7] = float("nan")
        flat[4, :] *= 1e30 if dtype != "fp16" else 6e3
        flat[5, :] *= 1e-38 if dtype == "fp32" else 1e-7
        flat[6, 0:8] = 0.0
        xt = x.to(TD[dtype])
        bits = 8 if mode == "sym" else 4
        fn_o = O.sym_quant if mode == "sym" else O.asym_quant
        with np.errstate(all="ignore"):
            want = fn_o(xt.float().numpy(), dtype, bits, False)
        fn = SymQuantizer if mode == "sym" else AsymQuantizer
        got = fn.apply(xt.to(cuda), torch.tensor([-2.0, 2.0]), bits, False)
        assert bits_equal(to_np(got), want), (mode, dtype, shape)
