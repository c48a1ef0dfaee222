"""GPU: one QAT step on a tiny Llama (the restatement used by bench.py --workload qat): losses are
finite, every decoder linear is a QuantizeLinear, and the fused fake-quant matches the oracle
inside the module."""
import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O

pytestmark = pytest.mark.gpu


def test_tiny_qat_step(cuda):
    from mxq_b200 import qat, QuantizeLinear, ops
    cfg = qat.llama_config(layers=2, hidden=256, inter=704, heads=4, vocab=512, seqlen=128)
    student, teacher, nq = qat.build_models(cfg, cuda, dtype=torch.bfloat16)
    assert nq == 2 * 7
    assert isinstance(student.model.layers[0].mlp.down_proj, QuantizeLinear)
    assert type(student.lm_head) is torch.nn.Linear
    opt = torch.optim.AdamW(student.parameters(), lr=1e-4)
    ids = torch.randint(0, 512, (2, 128), device=cuda)
    w0 = student.model.layers[0].self_attn.q_proj.weight.detach().clone()
    l0 = float(qat.qat_step(student, teacher, ids, opt))
    l1 = float(qat.qat_step(student, teacher, ids, opt))
    assert np.isfinite(l0) and np.isfinite(l1)
    assert not torch.equal(w0, student.model.layers[0].self_attn.q_proj.weight.detach())   # weights train through the STE
    w = student.model.layers[1].mlp.up_proj.weight.detach()
    want = O.fakequant_fwd(w.float().cpu().numpy(), "bf16", 2)
    got = ops.fakequant_fwd(w).float().cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
