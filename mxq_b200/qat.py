"""QAT step restatement for the benchmark (BASELINE configs[3]).

The reference's training driver (LLM-QAT/train.py:44-151, utils/kd_trainer.py:53-127) is out of
scope and not importable here (apex/fairscale); what one optimisation step does is restated:
teacher forward under no_grad (kd_trainer.py:55-59), student forward with every decoder linear a
QuantizeLinear (modeling_llama_quant.py:210-230,262-291; lm_head stays nn.Linear, :795), the
KL(student || teacher) loss of kd_trainer.py:42-48, backward through per-layer gradient
checkpointing (run_train.sh:38 -> fake-quant runs twice per step) and AdamW (run_train.sh:24-38).
The Llama module itself is the installed `transformers` implementation (library code); only the
linears are swapped for mxq_b200.QuantizeLinear.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils_quant import QuantizeLinear, group_quantize_linears


def convert_linears(model: nn.Module, w_bits: int = 2, skip=("lm_head",), group: bool = True) -> int:
    """Replace nn.Linear modules (except `skip`) by QuantizeLinear sharing the same weight.  group=True:
    siblings of one width under one parent (q/k/v/o, gate/up) fake-quantize in one launch per forward."""
    n = 0
    for name, mod in list(model.named_modules()):
        converted = []
        for cname, child in list(mod.named_children()):
            full = f"{name}.{cname}" if name else cname
            if type(child) is nn.Linear and not any(s in full for s in skip):
                q = QuantizeLinear(child.in_features, child.out_features, w_bits=w_bits, a_bits=32)
                q.weight = child.weight
                setattr(mod, cname, q)
                converted.append(q)
                n += 1
        if group and len(converted) > 1:
            group_quantize_linears(converted)
    return n


def kd_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor) -> torch.Tensor:
    """kd_trainer.py:42-48: F.kl_div(log_softmax(student), softmax(teacher), 'batchmean')."""
    s = F.log_softmax(student_logits.float().view(-1, student_logits.shape[-1]), dim=-1)
    t = F.softmax(teacher_logits.float().view(-1, teacher_logits.shape[-1]), dim=-1)
    return F.kl_div(s, t, reduction="batchmean")


def llama_config(layers: int = 32, hidden: int = 4096, inter: int = 11008, heads: int = 32, vocab: int = 32000,
                 seqlen: int = 2048):
    from transformers import LlamaConfig
    return LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                       num_attention_heads=heads, num_key_value_heads=heads, vocab_size=vocab,
                       max_position_embeddings=seqlen, rms_norm_eps=1e-5, use_cache=False)


def build_models(cfg, device, dtype=torch.bfloat16, w_bits: int = 2, seed: int = 0):
    """Random-init student (QuantizeLinear, gradient checkpointing) and frozen teacher with the
    same initial weights (train.py:59-90)."""
    from transformers import LlamaForCausalLM
    torch.manual_seed(seed)
    with torch.device(device):
        student = LlamaForCausalLM(cfg).to(dtype)
        teacher = LlamaForCausalLM(cfg).to(dtype)
    teacher.load_state_dict(student.state_dict())
    teacher.eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    nq = convert_linears(student, w_bits=w_bits)
    student.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    student.train()
    return student, teacher, nq


def qat_step(student, teacher, input_ids, optimizer) -> torch.Tensor:
    with torch.no_grad():
        t_logits = teacher(input_ids=input_ids, use_cache=False).logits
    s_logits = student(input_ids=input_ids, use_cache=False).logits
    loss = kd_loss(s_logits, t_logits)
    loss.backward()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach()
