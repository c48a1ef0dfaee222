"""Drop-in for ``WrappedGPT`` (mxq_quant/lib/layerwrapper.py:5-35): the Wanda activation statistic
scaler_row[k] = mean over samples of ||X[:, k]||_2^2, computed by csrc/calib.cu."""
from __future__ import annotations

import torch

from . import ops


class WrappedGPT:
    def __init__(self, layer, layer_id=0, layer_name="none"):
        self.layer = layer
        self.dev = self.layer.weight.device
        self.rows = layer.weight.data.shape[0]
        self.columns = layer.weight.data.shape[1]
        self.scaler_row = torch.zeros((self.columns), device=self.dev)
        self.nsamples = 0
        self.layer_id = layer_id
        self.layer_name = layer_name

    def add_batch(self, inp, out=None):
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        inp = inp.reshape((-1, inp.shape[-1]))
        prev = self.nsamples / (self.nsamples + tmp)         # layerwrapper.py:31
        self.nsamples += tmp
        ops.colsumsq(inp, out=self.scaler_row, prev_scale=prev, add_scale=1.0 / self.nsamples)

    def metric(self, W=None):
        """|W| * sqrt(scaler_row) (mxq_quant/lib/prune.py:177)."""
        W = self.layer.weight.data if W is None else W
        return ops.wanda_metric(W, self.scaler_row)
