"""Drop-in for the hot-path symbols of LLM-QAT/models/utils_quant.py:
``MXAsymQuantizer`` (:310-475) and ``QuantizeLinear`` (:601-727), backed by the fused sm_100a
kernels in csrc/fakequant.cu.  Same names, argument order and autograd contract as the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class MXAsymQuantizer(torch.autograd.Function):
    """Mixed 2/4-bit min-max fake quantization with a clipped straight-through backward.

    forward(input[N, K], clip_val[2], num_bits, layerwise) -> Tensor[N, K]   (utils_quant.py:315-462)
    backward -> (grad_input, None, None, None)                              (utils_quant.py:464-475)
    """

    @staticmethod
    def forward(ctx, input, clip_val, num_bits, layerwise):
        ctx.save_for_backward(input, clip_val)
        if layerwise:
            # utils_quant.py:334-336 sets alpha/beta but never `s`, so the reference raises
            # NameError at :458 -- the branch is dead; fail loudly instead of inventing semantics.
            raise NotImplementedError("MXAsymQuantizer(layerwise=True) is undefined in the reference")
        if input.dim() != 2:
            # :337 also admits 3-D tensors but slices dim 1 with column indices (an indexing
            # accident); 4-D leaves `s` undefined.  Weights are always 2-D (:630).
            raise NotImplementedError("MXAsymQuantizer is implemented for 2-D weight tensors")
        return ops.fakequant_fwd(input, num_bits=int(num_bits), group=16)

    @staticmethod
    def backward(ctx, grad_output):
        input, clip_val = ctx.saved_tensors
        lo, hi = (float(v) for v in clip_val.detach().to("cpu", torch.float32).tolist())
        if input.dtype != torch.float32:
            # input.ge(clip_val[1]) compares in the tensor dtype: the 0-dim clip value is cast
            lo = float(torch.tensor(lo).to(input.dtype))
            hi = float(torch.tensor(hi).to(input.dtype))
        grad_input = ops.ste_bwd(grad_output, input, lo, hi)
        return grad_input, None, None, None


class QuantizeLinear(nn.Linear):
    """nn.Linear whose weight is fake-quantized on every forward (utils_quant.py:601-727).
    Constructor keywords and state-dict keys (``weight`` only; bias always off, :613) match."""

    def __init__(self, *kargs, symmetric=True, bias=False, w_bits=32, a_bits=32,
                 act_layerwise=False, weight_layerwise=False, is_qk=False):
        super().__init__(*kargs, bias=False)
        self.w_bits = w_bits
        self.a_bits = a_bits
        self.act_layerwise = act_layerwise
        self.weight_layerwise = weight_layerwise
        self.is_qk = is_qk
        self.symmetric = symmetric
        if self.a_bits < 32 and self.a_bits > 2:
            # Sym/AsymQuantizer activation quantizers (utils_quant.py:31-199) are the "next" row
            # of SURVEY.md 8f; they are off in the reference recipe `run_train.sh 2 32 32`.
            raise NotImplementedError("activation quantization (a_bits < 32) is not on the MXQ hot path yet")

    def forward(self, input_):
        assert len(self.weight.size()) == 2
        real_weights = self.weight
        if self.w_bits >= 32:
            weight = self.weight
        elif self.w_bits >= 2:
            weight_clip_val = torch.tensor([-2.0, 2.0])          # utils_quant.py:636
            weight = MXAsymQuantizer.apply(real_weights, weight_clip_val, self.w_bits,
                                           self.weight_layerwise)
        else:
            # w_bits == 1 sign quantizer / BiT-style branch (:649-715): outside the MXQ path.
            raise NotImplementedError("w_bits < 2 is not part of the MXQ hot path")
        out = nn.functional.linear(input_, weight)
        if self.bias is not None:
            out += self.bias.view(1, -1).expand_as(out)
        return out
