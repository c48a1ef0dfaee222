"""Drop-in for ``MXQGPT`` (mxq_quant/lib/mxqgpt.py:353-452), the live PTQ layer quantizer behind
``--prune_method mxq``.

The reference accumulates a K x K fp32 Hessian per linear (68.7 GFLOP per 2048-token sample at
K=4096) and then only reads ``diag(H) == 0`` (mxqgpt.py:399-403).  Here ``add_batch`` keeps the
diagonal only (one HBM-bound pass over the activations, csrc/calib.cu) and ``fasterquant`` is one
fused kernel chain (csrc/ptq.cu) producing bit-identical fp16 fake-quant weights.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class MXQGPT:
    def __init__(self, layer):
        if not isinstance(layer, nn.Linear):
            # Conv2d / transformers.Conv1D flatten/transposes (mxqgpt.py:359-362) are never hit
            # by the Llama path.
            raise NotImplementedError("MXQGPT supports nn.Linear layers")
        self.layer = layer
        self.dev = self.layer.weight.device
        self.rows = layer.weight.shape[0]
        self.columns = layer.weight.shape[1]
        # diag(H) of the reference's H (mxqgpt.py:365,377-383)
        self.diagH = torch.zeros(self.columns, dtype=torch.float32, device=self.dev)
        self.nsamples = 0
        self.save_quant_dict = {}
        self.packed = None
        self._ws = None

    @property
    def H(self):
        raise AttributeError("the K x K Hessian is never materialised; use .diagH (its diagonal)")

    def add_batch(self, inp, out=None):
        """mxqgpt.py:369-383: H = H*n/(n+b) + (2/(n+b)) X^T X, restricted to the diagonal."""
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        inp = inp.reshape((-1, inp.shape[-1]))
        prev = self.nsamples / (self.nsamples + tmp)
        self.nsamples += tmp
        need = ops.L.lib().mxq_colsumsq_workspace_bytes(inp.shape[0], inp.shape[1])
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=inp.device)
        ops.colsumsq(inp, out=self.diagH, prev_scale=prev, add_scale=2.0 / self.nsamples,
                     workspace=self._ws)

    def fasterquant(self, blocksize=128, percdamp=.01, pack=False):
        """mxqgpt.py:387-448.  ``percdamp`` is unused by the reference too.  ``pack=True`` (an
        extension) additionally packs the ORIGINAL weights into the mixed 2/4-bit layout
        (``self.packed``) before they are replaced by their fake-quantized values."""
        # mxqgpt.py:413-415: the 48 two-bit columns of every 64-column block are cut into groups
        # range(0, 48, blocksize): 16 (what nas_quant passes, prune.py:409), 32 -> 32 + 16, and
        # anything >= 48 (the signature's default 128) -> one 48-wide group.
        width = min(int(blocksize), 48)
        if width not in (16, 32, 48):
            raise NotImplementedError("fasterquant: blocksize must be 16, 32 or >= 48")
        if pack and width != 16:
            raise ValueError("the packed layout has 16-column groups (gemv_mxq_cuda.cu:131-136): pack=True needs blocksize=16")
        W = self.layer.weight.data
        if W.dtype != torch.float16:
            raise TypeError("MXQGPT.fasterquant expects an fp16 layer (main.py loads the model in fp16)")
        # nsamples == 0: H is all zero, so every column is "dead" and W becomes 0 (mxqgpt.py:399-403);
        # the zero-initialised diagonal reproduces that
        colstat = self.diagH
        if pack:
            # self.packed may hold pre-allocated output tensors (nas_quant carves them from one arena)
            Wq, self.packed = ops.ptq_quant_pack(W, colstat, packed=self.packed)
        else:
            Wq = ops.ptq_quant(W, colstat, low_bits=2, group=width)
        self.layer.weight.data = Wq.reshape(self.layer.weight.shape).to(self.layer.weight.data.dtype)

    def free(self):
        # (the reference also calls torch.cuda.empty_cache() here, mxqgpt.py:452: a device-wide
        # synchronisation per linear that only matters for its 64-460 MB Hessians)
        self.diagH = None
        self._ws = None
