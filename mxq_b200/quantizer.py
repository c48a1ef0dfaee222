"""Drop-in for the path-relevant part of ``Quantizer`` (mxq_quant/lib/quantizer.py:23-180):
bits 2..8, perchannel, asymmetric, float zero-point, optional second-level 4-bit quantisation of
the scales over 16 consecutive rows.  Other configurations raise."""
from __future__ import annotations

import torch

from . import ops


class Quantizer:
    def __init__(self, shape=1):
        self.maxq = torch.tensor(0)
        self.scale = torch.zeros(shape)
        self.zero = torch.zeros(shape)

    def configure(self, bits, perchannel=False, sym=True, norm=2.0, grid=100, maxshrink=0.8,
                  round_zero: bool = False, qq_scale_bits=None, qq_zero_bits=None, qq_groupsize=16,
                  qq_zero_sym=False, reserved_bins: int = 0, qqq_params=None):
        if not perchannel or sym or round_zero or qq_zero_bits is not None or reserved_bins or \
                qq_groupsize != 16 or bits < 2:
            raise NotImplementedError(
                "only the mxq configuration is implemented: perchannel=True, sym=False, "
                "round_zero=False, qq_zero_bits=None, qq_groupsize=16 (mxqgpt.py:421,434)")
        self.bits = bits
        self.maxq = torch.tensor(2 ** bits - 1)
        self.qq_scale_bits = qq_scale_bits

    def find_params(self, x, weight=False):
        if not weight or x.dim() != 2:
            raise NotImplementedError("find_params is implemented for 2-D weights (weight=True)")
        self._y, self._codes, scale, zero = ops.rowquant(x, self.bits, self.qq_scale_bits)
        self._x = x
        self.scale = scale.reshape(-1, 1)
        self.zero = zero.reshape(-1, 1)

    def _run(self, x):
        if x is self._x:
            return self._y, self._codes
        raise NotImplementedError("quantize_dequantize is only supported on the tensor passed to find_params")

    def quantize_dequantize(self, x):
        return self._run(x)[0]

    def quantize(self, x):
        return self._run(x)[1].float()

    def dequantize(self, q):
        return self.scale * (q - self.zero)

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)
