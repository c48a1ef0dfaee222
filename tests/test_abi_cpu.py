"""CPU: the C-ABI library loads without a GPU, exports every symbol include/mxq_b200.h declares,
and rejects bad arguments before touching the device; the host mirrors fail loudly off-GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from mxq_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mxq_b200.h")).read()
    return sorted(set(re.findall(r"MXQ_API[^;(]*?\b(mxq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol():
    L = _lib.lib()
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert L.mxq_version() == 100
    assert b"shape" in L.mxq_error_string(-2)


def test_argument_errors_without_gpu():
    L = _lib.lib()
    buf = (C.c_char * 4096)()
    p = C.addressof(buf)
    p16 = (p + 15) & ~15
    # NULL pointer, unknown dtype, bad divisibility, misaligned pointer: all rejected up front
    assert L.mxq_fakequant_fwd(None, p16, None, 4, 64, 0, 16, 2, None, None) == -1
    assert L.mxq_fakequant_fwd(p16, p16, None, 4, 64, 7, 16, 2, None, None) == -3
    assert L.mxq_fakequant_fwd(p16, p16, None, 4, 48, 0, 16, 2, None, None) == -2
    assert L.mxq_fakequant_fwd(p16 + 4, p16, None, 4, 64, 0, 16, 2, None, None) == -4
    assert L.mxq_fakequant_fwd(p16, p16, None, 4, 64, 0, 24, 2, None, None) == -2   # group not 2^k
    assert L.mxq_fakequant_fwd(p16, p16, None, 0, 64, 0, 16, 2, None, None) == 0    # empty is a no-op
    assert L.mxq_ste_bwd(None, p16, p16, 16, 0, -2.0, 2.0, None) == -1
    assert L.mxq_ste_bwd(p16, p16, p16, 16, 9, -2.0, 2.0, None) == -3
    assert L.mxq_ptq_quant(p16, p16, None, None, 8, 64, 16, 2, None, p16, 4096, None) == -2   # rows % 16
    assert L.mxq_ptq_quant(p16, p16, None, None, 16, 64, 24, 2, None, p16, 4096, None) == -5  # group not 16 / 32 / 48
    # the mask-driven path needs its workspace (the reference recipe runs one fused kernel without any)
    assert L.mxq_ptq_quant(p16, p16, None, None, 16, 64, 16, 2, p16, p16, 8, None) == -6      # workspace
    assert L.mxq_rowquant(p16, None, None, None, None, 8, 16, 2, 4, None) == -2               # rows % 16 w/ qq
    pk = _lib.PackedC(p16, p16, p16, p16, p16, p16, p16)
    assert L.mxq_gemv(p16, pk, p16, 1, 100, 64, None) == -2
    assert L.mxq_pack(p16, None, 8, 64, pk, p16, 4096, None) == -2                            # OC % 16
    assert L.mxq_awq_gemv(p16, p16, p16, p16, p16, 1, 128, 8, 48, None) == -5
    # persistent decode chain: the plan is host code; every argument error is reported before a driver is needed
    nplan = L.mxq_gemv_chain_plan_bytes()
    assert nplan > 64 * 128
    plan = (C.c_char * (nplan + 64))()
    pp = (C.addressof(plan) + 63) & ~63
    job = _lib.GemvJobC(p16, p16, pk, 4096, 4096, -1, 0)
    jobs = (_lib.GemvJobC * 2)(job, job)
    assert L.mxq_gemv_chain_plan(None, 1, pp) == -1
    assert L.mxq_gemv_chain_plan(jobs, 1, pp + 8) == -4                                       # plan not 64-byte aligned
    assert L.mxq_gemv_chain_plan(jobs, 0, pp) == -2 and L.mxq_gemv_chain_plan(jobs, 65, pp) == -2
    jobs[0].IC = 4096 + 64                                                                    # IC % 256
    assert L.mxq_gemv_chain_plan(jobs, 1, pp) == -5
    jobs[0].IC, jobs[0].OC = 4096, 4096 + 8                                                   # OC % 32
    assert L.mxq_gemv_chain_plan(jobs, 1, pp) == -5
    jobs[0].OC, jobs[1].dep = 4096, 1                                                         # dep must name an EARLIER job
    assert L.mxq_gemv_chain_plan(jobs, 2, pp) in (-2, -5)                                     # (-5 first on a host without a driver)
    jobs[0].x = None
    assert L.mxq_gemv_chain_plan(jobs, 1, pp) == -1
    assert L.mxq_gemv_chain_run(None, pp, pp, 0, None) == -1
    assert L.mxq_ptq_workspace_bytes(4096, 4096) >= 4096 + 4096 * 8
    assert L.mxq_colsumsq_workspace_bytes(262144, 4096) > 0


def test_no_cpu_fallback():
    x = torch.zeros(4, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fakequant_fwd(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ste_bwd(x, x, -2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.colsumsq(x)
    from mxq_b200 import MXAsymQuantizer
    with pytest.raises(RuntimeError, match="CUDA"):
        MXAsymQuantizer.apply(x, torch.tensor([-2.0, 2.0]), 2, False)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmxq_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mxq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_reference_recipe_mask():
    gb = ops.reference_group_bits(128, 16, 2)
    assert gb.tolist() == [2, 2, 2, 0x84] * 2
    with pytest.raises(ValueError):
        ops.reference_group_bits(96, 16, 2)


def test_mirror_signatures_match_reference():
    import inspect
    from mxq_b200 import QuantizeLinear, MXQGPT, WrappedGPT
    from mxq_b200 import engine
    sig = inspect.signature(QuantizeLinear.__init__)
    for kw, default in dict(symmetric=True, bias=False, w_bits=32, a_bits=32, act_layerwise=False,
                            weight_layerwise=False, is_qk=False).items():
        assert sig.parameters[kw].default == default
    assert list(inspect.signature(MXQGPT.fasterquant).parameters)[:3] == ["self", "blocksize", "percdamp"]
    assert inspect.signature(MXQGPT.fasterquant).parameters["blocksize"].default == 128
    assert list(inspect.signature(MXQGPT.add_batch).parameters)[:3] == ["self", "inp", "out"]
    assert list(inspect.signature(WrappedGPT.__init__).parameters) == ["self", "layer", "layer_id", "layer_name"]
    assert list(inspect.signature(engine.gemv_mxq_forward_cuda).parameters) == [
        "in_feats", "kernel", "kernel_last", "zeros_and_scales", "scales_2nd", "zeros_2nd",
        "scales_4b", "zeros_4b", "group_size"]
    lin = QuantizeLinear(64, 32, w_bits=2)
    assert list(lin.state_dict().keys()) == ["weight"]
    from mxq_b200 import AsymQuantizer, SymQuantizer
    assert QuantizeLinear(64, 32, w_bits=2, a_bits=8).act_quantizer is SymQuantizer      # :622-626
    assert QuantizeLinear(64, 32, w_bits=2, a_bits=8, symmetric=False).act_quantizer is AsymQuantizer
    assert not hasattr(QuantizeLinear(64, 32, w_bits=2, a_bits=2), "act_quantizer")
    with pytest.raises(RuntimeError, match="CUDA"):                 # no CPU fallback
        SymQuantizer.apply(torch.zeros(2, 128), torch.tensor([-2.0, 2.0]), 8, False)


def test_segquant_plan_follows_reference_slicing():
    # utils_quant.py:56-64 / :144-157: 2-D -> column groups; 3-D -> per-token statistics with
    # tokens beyond (C // G) * G dead; 4-D -> per (b, h); layerwise -> one segment
    assert ops.segquant_plan((16, 256), "sym", False) == (32, 128, 1, 1)
    assert ops.segquant_plan((16, 64), "asym", False) == (128, 8, 1, 1)
    assert ops.segquant_plan((2, 12, 256), "sym", False) == (24, 256, 1, 1)
    assert ops.segquant_plan((1, 130, 128), "sym", False) == (130, 128, 130, 128)
    assert ops.segquant_plan((2, 20, 16), "asym", False) == (40, 16, 20, 16)
    assert ops.segquant_plan((2, 3, 8, 16), "sym", False) == (6, 128, 1, 1)
    assert ops.segquant_plan((4, 256), "asym", True) == (1, 1024, 1, 1)


def test_gemm_tile_plan_host_arithmetic():
    """mxq_gemm_plan (no device needed): whole tiles + K-split tail tiles of the CTA-pair GEMM on a
    148-SM part -- the tail is cut so that its clusters fit one wave, slices keep >= 16 K blocks, and
    the workspace holds one fp32 accumulator slot per (tail tile, slice, CTA of the pair)."""
    import ctypes as C
    L = _lib.lib()

    def plan(M, IC, OC, sms=148):
        out = (C.c_int32 * 5)()
        ws = C.c_size_t(0)
        assert L.mxq_gemm_plan(M, IC, OC, sms, out, C.byref(ws)) == 0
        return list(out), ws.value

    slot = 256 * 256 * 4
    # Llama-2-7B at M = 2048: 4096^2 = 64 tiles (one partial wave, nothing to cut) ...
    assert plan(2048, 4096, 4096) == ([4, 16, 64, 64, 1], 256)
    # ... gate/up 11008 x 4096 = 172 tiles = 2 waves of 74 + 24 tail tiles x 3 K slices
    p, ws = plan(2048, 4096, 11008)
    assert p == [4, 43, 172, 148, 3] and ws == 256 + 24 * 3 * 2 * slot
    # short M: 16 tiles x 4 slices (K = 4096 allows at most 4 slices of 16 K blocks)
    assert plan(128, 4096, 4096)[0] == [1, 16, 16, 0, 4]
    assert plan(128, 8192, 1024)[0] == [1, 4, 4, 0, 8]          # capped at 8 slices
    assert plan(2048, 1024, 4096)[0][4] == 1                     # K too short to cut
    assert plan(2048, 4096 + 64, 11008)[0][4] == 1               # IC % 256 != 0: generic K path, no cut
    # 43 tiles (more than half a wave): cutting them 3 ways was measured slower than whole tiles
    assert plan(512, 4096, 11008)[0] == [1, 43, 43, 43, 1]
    assert plan(2048, 8192, 3584)[0][4] == 1                    # 56 tiles, K = 8192: nothing gained
    for M, IC, OC in ((2048, 8192, 28672), (512, 4096, 2560), (1000, 4096, 1304), (2048, 28672, 1024)):
        (mt, nt, tiles, full, split), ws = plan(M, IC, OC)
        tail = tiles - full
        assert tiles == mt * nt and 0 <= full <= tiles and 1 <= split <= 8
        if split > 1:
            assert tail * split <= 74 and full % 74 == 0 and (IC // 64) // split >= 16
            assert ws == 256 + tail * split * 2 * slot
        else:
            assert full == tiles and ws == 256
    # another part: 132 SMs
    assert plan(2048, 4096, 11008, sms=132)[0] == [4, 43, 172, 172, 1]     # 40 tail tiles on 66 pairs: no cut
    assert L.mxq_gemm_plan(0, 4096, 4096, 148, (C.c_int32 * 5)(), None) != 0
