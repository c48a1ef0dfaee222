"""GPU parity for the tcgen05/TMEM prefill dequant-GEMM: vs the fp64 dequantize-then-matmul oracle
within max|err| <= 1e-3 * max|y| (fp16 in/out, fp32 accumulate)."""
import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O
from tests.gpu_util import packed_to_dev

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _rel(got, ref):
    return float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max())


# M > 256 runs the CTA-pair (cta_group::2) kernel, M <= 256 the single-CTA one
@pytest.mark.parametrize("M,OC,IC", [(256, 256, 64), (256, 256, 256), (128, 512, 4096), (300, 264, 1024),
                                     (512, 256, 64), (512, 256, 256), (1024, 512, 1024), (700, 264, 4096)])
def test_dense_pipeline(cuda, M, OC, IC):
    """tcgen05 + TMA + TMEM pipeline alone (dense fp16 B operand through TMA)."""
    from mxq_b200 import ops
    torch.manual_seed(M + OC + IC)
    x = torch.randn(M, IC, device=cuda).half()
    W = (torch.randn(OC, IC, device=cuda) * 0.05).half()
    y = ops.gemm_dense(x, W)
    ref = (x.double() @ W.double().T).cpu().numpy()
    assert _rel(y.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("M,OC,IC", [(256, 256, 64), (256, 256, 4096), (2048, 512, 4096), (77, 264, 128),
                                     (512, 256, 11008), (1000, 1024, 8192), (512, 256, 64), (512, 256, 512),
                                     (257, 136, 192), (2048, 264, 320)])
def test_packed_random_bits(cuda, M, OC, IC):
    from mxq_b200 import ops
    p = O.random_packed(OC, IC, seed=M + OC + IC)
    rng = np.random.default_rng(M)
    x = rng.standard_normal((M, IC)).astype(np.float16)
    ref = O.gemm_mxq_f32(x, p)
    y = ops.gemm(torch.from_numpy(x).to(cuda), packed_to_dev(p, cuda)).cpu().numpy()
    assert y.shape == (M, OC)
    assert _rel(y, ref) <= TOL


def test_full_size_vs_unpack_matmul(cuda):
    """Llama-2-7B q_proj at M=2048: GEMM == cuBLAS matmul on the unpacked weights; GEMV rows agree."""
    from mxq_b200 import ops
    torch.manual_seed(0)
    W = (torch.randn(4096, 4096, device=cuda) * 0.02).half()
    p = ops.pack(W)
    x = torch.randn(2048, 4096, device=cuda).half()
    y = ops.gemm(x, p).float()
    ref = x.float() @ ops.unpack(p).T
    assert float((y - ref).abs().max() / ref.abs().max()) <= TOL
    yv = ops.gemv(x[:4].contiguous(), p).float()
    assert float((yv - y[:4]).abs().max() / ref.abs().max()) <= TOL


# K-split tail tiles (pair::Plan): tiles % SM-pairs != 0 and IC >= 2048 cut the last wave's tiles into
# K slices that meet in the fp32 exchange workspace
@pytest.mark.parametrize("M,OC,IC", [(512, 2560, 2048), (1000, 1304, 4096), (512, 520, 8192), (2048, 11008, 2048)])
def test_packed_split_k_tail(cuda, M, OC, IC):
    from mxq_b200 import ops
    assert ops.gemm_workspace_bytes(M, IC, OC) > 4096, "shape was meant to exercise the K split"
    p = O.random_packed(OC, IC, seed=M + OC + IC)
    rng = np.random.default_rng(M)
    x = rng.standard_normal((M, IC)).astype(np.float16)
    ref = O.gemm_mxq_f32(x, p)
    pd = packed_to_dev(p, cuda)
    xd = torch.from_numpy(x).to(cuda)
    ws = ops.gemm_workspace(M, IC, OC, cuda)
    ws.fill_(0xFF)                      # the workspace need not be initialised
    y1 = ops.gemm(xd, pd, workspace=ws)
    assert _rel(y1.cpu().numpy(), ref) <= TOL
    # slices are added in slice order: bit-identical from run to run
    for _ in range(3):
        y2 = ops.gemm(xd, pd, workspace=ws)
        assert torch.equal(y1, y2)
    y3 = ops.gemm(xd, pd, workspace=None)
    assert torch.equal(y1, y3)
    # guard bands: nothing is written past the declared workspace size or outside the output
    need = ops.gemm_workspace_bytes(M, IC, OC)
    wsg = torch.full((need + 8192,), 0xAB, dtype=torch.uint8, device=cuda)
    yg = torch.full((M + 2, OC), 7.0, dtype=torch.float16, device=cuda)
    from mxq_b200 import _lib as L
    rc = L.lib().mxq_gemm(L.ptr(xd), L.packed_struct(pd), yg[1:M + 1].data_ptr(), M, IC, OC, wsg.data_ptr(), need, L.stream())
    assert rc == 0
    assert torch.equal(yg[1:M + 1], y1)
    assert bool((wsg[need:] == 0xAB).all()) and bool((yg[0] == 7.0).all()) and bool((yg[M + 1] == 7.0).all())
    # a workspace that is too small falls back to whole tiles
    small = torch.empty(4096, dtype=torch.uint8, device=cuda)
    out = torch.empty_like(y1)
    rc = L.lib().mxq_gemm(L.ptr(xd), L.packed_struct(pd), L.ptr(out), M, IC, OC, L.ptr(small), small.numel(), L.stream())
    assert rc == 0
    assert _rel(out.cpu().numpy(), ref) <= TOL


def test_split_k_mlp_shape_vs_unpack_matmul(cuda):
    """Llama-2-7B gate/up shape at M = 2048 (172 tiles on 74 SM pairs: 24 tail tiles x 3 K slices)."""
    from mxq_b200 import ops
    torch.manual_seed(1)
    W = (torch.randn(11008, 4096, device=cuda) * 0.02).half()
    p = ops.pack(W)
    x = torch.randn(2048, 4096, device=cuda).half()
    y = ops.gemm(x, p).float()
    ref = x.float() @ ops.unpack(p).T
    assert float((y - ref).abs().max() / ref.abs().max()) <= TOL


def _outlier_x(M, IC, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((M, IC))
    ch = rng.choice(IC, 8, replace=False)
    x[:, ch] *= rng.uniform(20, 100, 8)
    return x.astype(np.float16)


@pytest.mark.parametrize("M,OC,IC", [(512, 512, 4096), (300, 264, 1024), (64, 512, 8192)])
def test_packed_outlier_activations_per_element(cuda, M, OC, IC):
    """north_star's tolerance per element, |y - y_ref| <= 1e-3 * sum_k |w_k x_k| (+ the fp16 store),
    with LLM-like outlier channels.  The GEMM dequantizes the weights to fp16 operands (one rounding
    of s * (q - z), 2^-11 relative) and accumulates in fp32 on the tensor cores."""
    from mxq_b200 import ops
    p = O.random_packed(OC, IC, seed=M + OC + IC)
    x = _outlier_x(M, IC, seed=IC)
    ref = O.gemm_mxq_f32(x, p)
    Wd = np.abs(O.decode_mxq(p).astype(np.float64))
    bound = 1e-3 * (np.abs(x.astype(np.float64)) @ Wd.T) + np.abs(ref) * 2.0 ** -11
    y = ops.gemm(torch.from_numpy(x).to(cuda), packed_to_dev(p, cuda)).cpu().numpy().astype(np.float64)
    err = np.abs(y - ref)
    assert (err <= bound).all(), f"worst err/bound {float((err / bound).max()):.3f}"


@pytest.mark.parametrize("OC,IC", [(4096, 11008), (3584, 8192), (1024, 28672)])
def test_full_size_more_shapes(cuda, OC, IC):
    """down_proj (4096 x 11008) and two Llama-2-70B shards at 8 ranks (gate/up 28672/8 x 8192, down_proj
    8192/8 x 28672) at M = 2048: GEMM == fp32 matmul on the unpacked weights, per element."""
    from mxq_b200 import ops
    torch.manual_seed(OC + IC)
    p = {}
    for k, (s, d) in ops.packed_shapes(OC, IC).items():
        if d == torch.float16:
            p[k] = (torch.rand(s, device=cuda) * 0.009 + 0.001).half()
        else:
            p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=cuda, dtype=torch.int64).to(torch.int32)
    x = torch.randn(2048, IC, device=cuda).half()
    x[:, 100:108] *= 20
    y = ops.gemm(x, p).float()
    Wd = ops.unpack(p)                                   # fp32 decode (bit-exact vs the oracle: test_gpu_packed)
    ref = x.float() @ Wd.T
    bound = 1e-3 * (x.float().abs() @ Wd.abs().T) + ref.abs() * 2.0 ** -11
    # the fp32 reference itself carries ~K * 2^-24 relative summation noise: well inside the bound
    assert bool(((y - ref).abs() <= bound).all()), float(((y - ref).abs() / bound).max())


def _awq_case(M, IC, OC, G, seed):
    rng = np.random.default_rng(seed)
    kernel = rng.integers(0, 2 ** 32, (IC, OC // 8), dtype=np.uint64).astype(np.uint32)
    zeros = rng.integers(0, 2 ** 32, (IC // G, OC // 8), dtype=np.uint64).astype(np.uint32)
    scales = rng.uniform(0.001, 0.01, (IC // G, OC)).astype(np.float16)
    x = rng.standard_normal((M, IC)).astype(np.float16)
    nib = np.array([0, 4, 1, 5, 2, 6, 3, 7], dtype=np.uint32)           # channel c of a word <- nibble nib[c]
    q = ((kernel[:, :, None] >> (4 * nib)) & 0xF).reshape(IC, OC).astype(np.float64)
    z = ((zeros[:, :, None] >> (4 * nib)) & 0xF).reshape(IC // G, OC).astype(np.float64)
    g = np.arange(IC) // G
    W = scales.astype(np.float64)[g] * (q - z[g])                        # [IC, OC]
    return x, kernel, scales, zeros, W


@pytest.mark.parametrize("M,IC,OC,G", [(48, 512, 256, 128), (512, 4096, 1024, 128), (300, 1024, 192, 64), (2048, 4096, 4096, 128)])
def test_awq_gemm_vs_oracle(cuda, M, IC, OC, G):
    """AWQ uniform 4-bit prefill GEMM (gemm_cuda_gen.cu:424-478) against the restated decode + fp64 matmul."""
    from mxq_b200 import engine
    x, kernel, scales, zeros, W = _awq_case(M, IC, OC, G, seed=M + OC)
    ref = x.astype(np.float64) @ W
    y = engine.gemm_forward_cuda(torch.from_numpy(x).to(cuda), torch.from_numpy(kernel.view(np.int32)).to(cuda),
                                 torch.from_numpy(scales).to(cuda), torch.from_numpy(zeros.view(np.int32)).to(cuda), 8)
    bound = 1e-3 * (np.abs(x.astype(np.float64)) @ np.abs(W)) + np.abs(ref) * 2.0 ** -11
    assert y.shape == (M, OC) and y.dtype == torch.float16
    assert (np.abs(y.cpu().numpy().astype(np.float64) - ref) <= bound).all()


def test_awq_gemm_matches_reference_kernel(cuda):
    """The reference's own gemm_forward_cuda (compiled from its sources into oracle/_ref/awq_gemm_ref.so by
    oracle/build_ref.py; the reference never builds it) on random bits: pins the nibble order of kernel and
    zeros and the scale indexing against real reference code.  The reference sums fp16 split-K partials, so
    the comparison is within fp16 accumulation noise of the fp64 oracle for both."""
    import importlib.util
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "awq_gemm_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/awq_gemm_ref.so not built (python oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("awq_gemm_ref", so)
    refmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refmod)
    from mxq_b200 import engine
    M, IC, OC, G = 16, 1024, 256, 128
    x, kernel, scales, zeros, W = _awq_case(M, IC, OC, G, seed=5)
    args = (torch.from_numpy(x).to(cuda), torch.from_numpy(kernel.view(np.int32)).to(cuda),
            torch.from_numpy(scales).to(cuda), torch.from_numpy(zeros.view(np.int32)).to(cuda))
    y_ref = refmod.gemm_forward_cuda(*args, 1).float().cpu().numpy()
    torch.cuda.synchronize()
    y = engine.gemm_forward_cuda(*args, 1).float().cpu().numpy()
    exact = x.astype(np.float64) @ W
    scale = np.abs(exact).max()
    assert np.abs(y_ref - exact).max() <= 4e-3 * scale, "the restated decode reproduces the reference kernel"
    assert np.abs(y - exact).max() <= 1e-3 * scale
    assert np.abs(y - y_ref).max() <= 4e-3 * scale
