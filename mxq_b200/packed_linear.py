"""Packed-checkpoint side of the MXQ path: the reference stops at fp16
fake-quantized weights (mxq_quant/main.py:96-100 saves them with ``save_pretrained``) and its only
consumer of the packed 2/4-bit layout is the raw ``gemv_mxq_forward_cuda`` binding
(cuda_kernel/test_correct_gemv.py:49).  This module closes the loop PTQ -> pack -> inference:

  MXQLinear            nn.Module holding the seven packed tensors of gemv_mxq_cuda.cu:39-208 as
                       buffers (state-dict keys = the binding's argument names); forward routes
                       decode-sized inputs to the GEMV kernel and prefill-sized ones to the
                       tcgen05 dequant-GEMM
  pack_linear          nn.Linear (fp16) -> MXQLinear, optionally with the calibration statistic
  decode_chain         several batch-1 MXQLinear calls as ONE persistent launch (ops.GemvChain)
  convert_model        swap every nn.Linear that nas_quant(args.pack=True) annotated
  save_packed / load_packed   the packed tensors of a converted model, as one torch file

There is no CPU fallback: MXQLinear.forward raises for non-CUDA inputs.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

PACKED_KEYS = ("weight", "weight_last", "zeros_and_scales", "zeros_2nd", "scales_2nd", "scales_4b", "zeros_4b")
FORMAT = "mxq_b200.packed.v1"


class MXQLinear(nn.Module):
    """y = x @ dequant(W)^T for a packed mixed 2/4-bit weight [out_features, in_features]."""

    GEMV_MAX_TOKENS = 8        # up to here the weight-streaming GEMV wins over a 256-token MMA tile

    def __init__(self, in_features: int, out_features: int, device=None, group_perm: torch.Tensor | None = None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        for k, (shape, dt) in ops.packed_shapes(out_features, in_features).items():
            self.register_buffer(k, torch.zeros(shape, dtype=dt, device=device))
        # importance-driven column order: packed 16-column group g = original group perm[g];
        # None = the reference's positional recipe
        self.register_buffer("group_perm", None if group_perm is None else group_perm.to(device=device, dtype=torch.int32))
        self._ws = None
        self._checked = False
        # programmatic dependent launch lets the GEMV fetch its packed weights while the previous
        # kernel of the stream still runs: only legal when that kernel does not write them, so it is
        # opt-in (a decode loop over resident weights sets `module.pdl = True`)
        self.pdl = False

    def _apply(self, fn, recurse=True):
        """nn.Module.to(dtype) / .half() / .bfloat16() / .float() cast floating buffers; the packed
        fp16 scale tensors are part of a bit-exact storage format, so only their DEVICE follows."""
        def keep_dtype(t):
            r = fn(t)
            return r if r.dtype == t.dtype else t.to(device=r.device)
        self._ws = None
        self._checked = False
        return super()._apply(keep_dtype, recurse)

    @property
    def packed(self) -> dict:
        return {k: getattr(self, k) for k in PACKED_KEYS}

    @classmethod
    def from_packed(cls, packed: dict, group_perm: torch.Tensor | None = None) -> "MXQLinear":
        OC, IC = packed["weight"].shape[0], packed["weight"].shape[1] * 16
        m = cls(IC, OC, device=packed["weight"].device, group_perm=group_perm)
        for k in PACKED_KEYS:
            getattr(m, k).copy_(packed[k])
        return m

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """The kernels compute in fp16 (fp32 accumulation) like the reference binding
        (gemv_mxq_cuda.cu:242-256); other input dtypes are cast in and the result is cast back."""
        if not x.is_cuda:
            raise RuntimeError("MXQLinear needs CUDA tensors (no CPU fallback)")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        if x2.dtype != torch.float16:
            x2 = x2.half()
        if x2.shape[0] == 0:
            return x.new_zeros((*lead, self.out_features))
        if not self._checked:
            ops._check_packed(self.packed)          # once per (re)materialisation of the buffers
            self._checked = True
        if x2.shape[0] <= self.GEMV_MAX_TOKENS:
            y = ops.gemv(x2, self.packed, validate=False, pdl=self.pdl, group_perm=self.group_perm)
        else:
            if self.group_perm is not None:
                x2 = ops.gather_groups(x2, self.group_perm)
            need = ops.gemm_workspace_bytes(x2.shape[0], self.in_features, self.out_features)
            if self._ws is None or self._ws.device != x2.device or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x2.device)
            y = ops.gemm(x2, self.packed, workspace=self._ws, validate=False)
        y = y.reshape(*lead, self.out_features)
        return y if x.dtype == torch.float16 else y.to(x.dtype)

    def dequantize(self, dtype=torch.float16) -> torch.Tensor:
        """The dequantized weight in the ORIGINAL column order."""
        W = ops.unpack(self.packed, dtype)
        if self.group_perm is None:
            return W
        inv = torch.empty_like(self.group_perm)
        inv[self.group_perm.long()] = torch.arange(self.group_perm.numel(), dtype=torch.int32, device=inv.device)
        return W.view(W.shape[0], -1, 16)[:, inv.long()].reshape(W.shape)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bits=3.0 (2/4 mixed)"


@torch.no_grad()
def pack_linear(linear: nn.Linear, colstat: torch.Tensor | None = None, importance: bool = False) -> MXQLinear:
    """Quantize an fp16 nn.Linear into the packed layout (encode policy: DESIGN.md section 2).
    `colstat`: calibration column statistic (diag(H) of MXQGPT / scaler_row of WrappedGPT), zero = dead
    column (mxqgpt.py:401-403).  importance=True: of every 4 groups the one with the largest Wanda
    metric |W| * sqrt(colstat) (prune.py:177) gets the 4 bits instead of the positionally last one."""
    if linear.bias is not None:
        raise ValueError("the MXQ path has no bias (utils_quant.py:613)")
    W = linear.weight.data
    if not W.is_cuda:
        raise RuntimeError("pack_linear needs CUDA weights (no CPU fallback)")
    W = W.half()
    perm = ops.importance_permutation(ops.allocate_group_bits(W, colstat)) if importance else None
    return MXQLinear.from_packed(ops.pack(W, colstat, group_perm=perm), group_perm=perm)


def decode_chain(jobs: list) -> "ops.GemvChain":
    """Batch-1 decode of several packed linears as ONE persistent launch (csrc/gemv_chain.cu).

    jobs: list of (MXQLinear, x, y, dep) -- x fp16 [in_features] (or [1, in_features]) and y fp16
    [out_features] are buffers the caller keeps and refills between runs; dep = -1 if x exists when the
    chain is launched, or the index of an earlier job whose y this x is (or is computed from by an earlier
    job).  q/k/v (one x), gate/up (one x), the experts of an MoE layer or the same linear of several
    sequences are the natural lists; a list may also chain linears on each other's outputs.  Returns the
    chain; ``chain.run()`` launches it on the current stream (capturable in a CUDA graph).  Linears with
    an importance permutation (``group_perm``) are not supported here: call them directly."""
    out = []
    for lin, x, y, dep in jobs:
        if not isinstance(lin, MXQLinear):
            raise TypeError("decode_chain takes MXQLinear modules")
        if lin.group_perm is not None:
            raise ValueError("decode_chain: linears with group_perm run through MXQLinear.forward")
        out.append((x, lin.packed, y, dep))
    return ops.GemvChain(out)


def convert_model(model: nn.Module) -> nn.Module:
    """Replace every nn.Linear carrying ``mxq_packed`` (set by nas_quant with args.pack) in place."""
    for name, child in list(model.named_children()):
        if isinstance(child, nn.Linear) and hasattr(child, "mxq_packed"):
            setattr(model, name, MXQLinear.from_packed(child.mxq_packed))
        else:
            convert_model(child)
    return model


def save_packed(model: nn.Module, path: str) -> None:
    """{"format", "linears": {module name: {packed key: tensor}}} of every MXQLinear in `model`."""
    linears = {n: dict({k: v.detach().cpu() for k, v in m.packed.items()},
                       **({"group_perm": m.group_perm.detach().cpu()} if m.group_perm is not None else {}))
               for n, m in model.named_modules() if isinstance(m, MXQLinear)}
    torch.save({"format": FORMAT, "linears": linears}, path)


def load_packed(model: nn.Module, path: str, device=None) -> nn.Module:
    """Swap the named nn.Linear / MXQLinear modules of `model` for MXQLinear loaded from `path`."""
    blob = torch.load(path, map_location="cpu")
    if blob.get("format") != FORMAT:
        raise ValueError(f"{path}: not an {FORMAT} file")
    for name, packed in blob["linears"].items():
        parent = model
        *path_, leaf = name.split(".")
        for part in path_:
            parent = getattr(parent, part)
        old = getattr(parent, leaf)
        dev = device
        if dev is None:
            t = next(iter(old.parameters()), None)
            if t is None:
                t = next(iter(old.buffers()))
            dev = t.device
        perm = packed.get("group_perm")
        setattr(parent, leaf, MXQLinear.from_packed({k: packed[k].to(dev) for k in PACKED_KEYS},
                                                    group_perm=None if perm is None else perm.to(dev)))
    return model
