"""Drop-in for the hot-path symbols of LLM-QAT/models/utils_quant.py:
``MXAsymQuantizer`` (:310-475), ``QuantizeLinear`` (:601-727) and the activation / KV-cache
quantizers ``SymQuantizer`` (:31-95) and ``AsymQuantizer`` (:98-199), backed by the fused sm_100a
kernels in csrc/fakequant.cu and csrc/actquant.cu.  Same names, argument order and autograd
contract as the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


# QuantizeLinear passes the constant [-2, 2] (utils_quant.py:636,719) on every forward: one shared
# CPU tensor and its per-dtype bounds are cached instead of being rebuilt 672 times per QAT step
_CLIP = torch.tensor([-2.0, 2.0])
_CLIP_CACHE: dict = {}


def _clip_bounds(clip_val, dtype):
    if clip_val is _CLIP:
        hit = _CLIP_CACHE.get(dtype)
        if hit is None:
            hit = _CLIP_CACHE[dtype] = _clip_bounds(_CLIP.clone(), dtype)
        return hit
    lo, hi = (float(v) for v in clip_val.detach().to("cpu", torch.float32).tolist())
    if dtype != torch.float32:
        # input.ge(clip_val[1]) compares in the tensor dtype: the 0-dim clip value is cast
        lo = float(torch.tensor(lo).to(dtype))
        hi = float(torch.tensor(hi).to(dtype))
    return lo, hi


class _SegQuantizer(torch.autograd.Function):
    """Shared body of SymQuantizer / AsymQuantizer: one statistic per contiguous segment
    (ops.segquant_plan states how the reference segments 2-D / 3-D / 4-D tensors), clipped
    straight-through backward (utils_quant.py:89-95,192-199)."""
    MODE = "sym"

    @classmethod
    def _fwd(cls, ctx, input, clip_val, num_bits, layerwise):
        ctx.save_for_backward(input, clip_val)
        if input.dim() > 4 or input.dim() < 2:
            raise ValueError
        nseg, seglen, period, valid = ops.segquant_plan(tuple(input.shape), cls.MODE, bool(layerwise))
        return ops.segquant_fwd(input, cls.MODE, int(num_bits), nseg, seglen, period, valid)

    @staticmethod
    def _bwd(ctx, grad_output):
        input, clip_val = ctx.saved_tensors
        lo, hi = _clip_bounds(clip_val, input.dtype)
        return ops.ste_bwd(grad_output, input, lo, hi), None, None, None


class SymQuantizer(_SegQuantizer):
    """|x|max uniform quantization, group 128 (utils_quant.py:31-95)."""
    MODE = "sym"

    @staticmethod
    def forward(ctx, input, clip_val, num_bits, layerwise):
        return SymQuantizer._fwd(ctx, input, clip_val, num_bits, layerwise)

    @staticmethod
    def backward(ctx, grad_output):
        return _SegQuantizer._bwd(ctx, grad_output)


class AsymQuantizer(_SegQuantizer):
    """min-max quantization, group 8 (utils_quant.py:98-199)."""
    MODE = "asym"

    @staticmethod
    def forward(ctx, input, clip_val, num_bits, layerwise):
        return AsymQuantizer._fwd(ctx, input, clip_val, num_bits, layerwise)

    @staticmethod
    def backward(ctx, grad_output):
        return _SegQuantizer._bwd(ctx, grad_output)


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
        lo, hi = _clip_bounds(clip_val, input.dtype)
        grad_input = ops.ste_bwd(grad_output, input, lo, hi)
        return grad_input, None, None, None


class MXAsymQuantizerMulti(torch.autograd.Function):
    """MXAsymQuantizer over several weights of one width in ONE launch each way it matters: forward =
    mxq_fakequant_fwd_multi, backward = the clipped straight-through estimator per tensor.  Numerically
    identical to calling MXAsymQuantizer.apply on every weight."""

    @staticmethod
    def forward(ctx, clip_val, num_bits, *weights):
        ctx.save_for_backward(clip_val, *weights)
        return tuple(ops.fakequant_fwd_multi(list(weights), num_bits=int(num_bits), group=16))

    @staticmethod
    def backward(ctx, *grads):
        clip_val, *weights = ctx.saved_tensors
        out = []
        for g, w in zip(grads, weights):
            if g is None:
                out.append(None)
            else:
                lo, hi = _clip_bounds(clip_val, w.dtype)
                out.append(ops.ste_bwd(g, w, lo, hi))
        return (None, None, *out)


class FakeQuantGroup:
    """QuantizeLinear modules of one width that are evaluated together in every forward (q/k/v/o of an
    attention block, gate/up of an MLP): the first member to run fake-quantizes ALL members' weights in
    one launch, the others pick their result up.  A result is handed out once per forward; the group
    falls back to per-module launches whenever its bookkeeping does not match (a member skipped, grad
    mode changed, weights updated in between)."""

    def __init__(self, members):
        self.members = list(members)
        for m in self.members:
            m._fq_group = self
        self._pending = {}
        self._key = None

    def weight_for(self, module):
        ws = [m.weight for m in self.members]
        key = (torch.is_grad_enabled(), tuple(w._version for w in ws), tuple(w.data_ptr() for w in ws))
        if id(module) not in self._pending or key != self._key:
            outs = MXAsymQuantizerMulti.apply(_CLIP, module.w_bits, *ws)
            self._pending = {id(m): o for m, o in zip(self.members, outs)}
            self._key = key
        return self._pending.pop(id(module))


def group_quantize_linears(modules) -> list:
    """Form FakeQuantGroups from QuantizeLinear siblings of equal in_features / w_bits / dtype (same parent
    module = evaluated in the same forward).  Returns the groups; un-groupable modules stay as they are."""
    buckets = {}
    for m in modules:
        if isinstance(m, QuantizeLinear) and 2 <= m.w_bits < 32 and not m.weight_layerwise:
            buckets.setdefault((m.in_features, m.w_bits, m.weight.dtype, m.weight.device), []).append(m)
    return [FakeQuantGroup(ms) for ms in buckets.values() if len(ms) > 1]


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
        if self.a_bits < 32 and self.a_bits > 2:                  # utils_quant.py:622-626
            self.act_quantizer = SymQuantizer if symmetric else AsymQuantizer

    def forward(self, input_):
        assert len(self.weight.size()) == 2
        real_weights = self.weight
        if self.w_bits >= 32:
            weight = self.weight
        elif self.w_bits >= 2:
            grp = getattr(self, "_fq_group", None)
            if grp is not None:
                weight = grp.weight_for(self)                    # one launch for the whole sibling group
            else:
                weight_clip_val = _CLIP                          # utils_quant.py:636
                weight = MXAsymQuantizer.apply(real_weights, weight_clip_val, self.w_bits,
                                               self.weight_layerwise)
        else:
            # w_bits == 1 sign quantizer / BiT-style branch (:649-715): outside the MXQ path.
            raise NotImplementedError("w_bits < 2 is not part of the MXQ hot path")
        if self.a_bits < 32 and self.a_bits > 2:                  # utils_quant.py:717-721
            act_clip_val = _CLIP
            input_ = self.act_quantizer.apply(input_, act_clip_val, self.a_bits, self.act_layerwise)
        out = nn.functional.linear(input_, weight)
        if self.bias is not None:
            out += self.bias.view(1, -1).expand_as(out)
        return out
