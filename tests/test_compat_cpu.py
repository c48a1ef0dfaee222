"""compat/: the reference's import names resolve to mxq_b200 (and, with MXQ_REFERENCE_ROOT, the
reference's own model file imports OUR QuantizeLinear unedited); host-side module behaviour that needs
no GPU: MXQLinear keeps its packed storage dtypes under .to(dtype), the C-ABI guard rejects CPU and
mixed-device tensors."""
import importlib
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "compat")
REF = "/root/reference"


def _run(code, env_extra=None):
    env = dict(os.environ, PYTHONPATH=COMPAT + os.pathsep + ROOT)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_import_names_resolve_to_mxq_b200():
    out = _run("import mxq_inference_engine as e, models.utils_quant as u, lib.mxqgpt as g, lib.prune as p, "
               "lib.quantizer as q, lib.layerwrapper as w\n"
               "print(e.gemv_mxq_forward_cuda.__module__, u.QuantizeLinear.__module__, u.MXAsymQuantizer.__module__, "
               "g.MXQGPT.__module__, p.nas_quant.__module__, q.Quantizer.__module__, w.WrappedGPT.__module__)\n"
               "print(sorted(n for n in ('prune_wanda','prune_magnitude','prune_sparsegpt','check_sparsity','find_layers','nas_quant') if hasattr(p, n)))")
    mods = out.split("\n")[0].split()
    assert all(m.startswith("mxq_b200.") for m in mods), mods
    assert "nas_quant" in out and "prune_wanda" in out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_model_file_picks_up_our_quantize_linear():
    """LLM-QAT/models/modeling_llama_quant.py:34 `from models.utils_quant import QuantizeLinear`, unedited."""
    out = _run("import models.modeling_llama_quant as m, models.utils_quant as u\n"
               "print(m.QuantizeLinear is u.QuantizeLinear, u.QuantizeLinear.__module__, m.__file__)",
               {"MXQ_REFERENCE_ROOT": REF})
    ok, mod, path = out.split()
    assert ok == "True" and mod == "mxq_b200.utils_quant" and path.startswith(REF)


def test_mxqlinear_keeps_packed_dtypes():
    from mxq_b200.packed_linear import MXQLinear
    m = MXQLinear(256, 64)
    m.scales_2nd.fill_(0.0123)
    ref = m.scales_2nd.clone()
    for cast in (lambda x: x.to(torch.bfloat16), lambda x: x.float(), lambda x: x.double(), lambda x: x.half()):
        m = cast(m)
        assert m.scales_2nd.dtype == torch.float16 and m.scales_4b.dtype == torch.float16
        assert m.weight.dtype == torch.int32 and m.zeros_4b.dtype == torch.int32
        assert torch.equal(m.scales_2nd, ref)           # no lossy round trip through bf16
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 256))


def test_abi_guard_rejects_cpu_tensors():
    from mxq_b200 import _lib as L
    with pytest.raises(RuntimeError, match="CUDA"):
        L.require_cuda(torch.zeros(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        L.on(torch.zeros(4))
    L.require_cuda(None)                                 # optional tensors may be absent


def test_quantize_linear_reuses_one_clip_tensor():
    from mxq_b200 import utils_quant as u
    lo, hi = u._clip_bounds(u._CLIP, torch.bfloat16)
    assert (lo, hi) == (-2.0, 2.0) and u._clip_bounds(u._CLIP, torch.bfloat16) is u._CLIP_CACHE[torch.bfloat16]
    assert u._clip_bounds(torch.tensor([-1.5, 0.3]), torch.float32) == (-1.5, 0.30000001192092896)
