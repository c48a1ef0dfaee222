"""Tensor-level wrappers over the C ABI (include/mxq_b200.h).  torch is plumbing here: it owns the
device memory and the stream; every computation happens in libmxq_b200.so."""
from __future__ import annotations

import torch

from . import _lib as L

POOL = L.POOL


def reference_group_bits(cols: int, group: int = 16, low_bits: int = 2, device=None) -> torch.Tensor:
    """{low, low, low, POOL|4} repeated -- the reference's positional 2/4-bit recipe
    (LLM-QAT/models/utils_quant.py:340-385; mxq_quant/lib/mxqgpt.py:404-419)."""
    if cols % (4 * group):
        raise ValueError(f"cols={cols} must be a multiple of 4*group={4 * group}")
    gb = torch.full((cols // group,), low_bits, dtype=torch.uint8)
    gb[3::4] = POOL | 4
    return gb.to(device) if device is not None else gb


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def fakequant_fwd(x: torch.Tensor, num_bits: int = 2, group: int = 16, group_bits=None,
                  return_codes: bool = False, uniform_low: bool = False):
    """MXAsymQuantizer.forward (utils_quant.py:315-462) for a 2-D tensor.  uniform_low=True promises
    that `group_bits` only MOVES the pooled 4-bit group (every other group is `num_bits` wide -- what
    allocate_group_bits produces): the row-resident kernel then serves the mask instead of the
    shared-memory ring kernel (same bits, 1.5-2x faster)."""
    if x.dim() != 2:
        raise ValueError("MXAsymQuantizer fake-quant is defined for 2-D weights (utils_quant.py:630)")
    L.require_cuda(x, group_bits)
    x = x.contiguous()
    rows, cols = x.shape
    if uniform_low and group_bits is not None and not return_codes and group in (16, 128) and cols % (4 * group) == 0 \
            and cols * x.element_size() <= 6144 * 16 and (x.element_size() == 4 or num_bits == 2):
        return fakequant_fwd_multi([x], num_bits, group, pooled_mask=group_bits)[0]
    if cols % 64 and group_bits is None and group == 16:
        # the reference itself fails here: W_4b has K/64*16 columns (utils_quant.py:347,368)
        raise ValueError("in_features must be a multiple of 64 for the mixed 2/4-bit recipe")
    out = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if return_codes else None
    with L.on(x) as st:
        rc = L.lib().mxq_fakequant_fwd(L.ptr(x), L.ptr(out), L.ptr(codes), rows, cols, L.dtype_enum(x),
                                       group, num_bits, L.ptr(group_bits), st)
    L.check(rc, "mxq_fakequant_fwd")
    return (out, codes) if return_codes else out


def fakequant_fwd_multi(xs: list, num_bits: int = 2, group: int = 16, pooled_mask: torch.Tensor | None = None,
                        outs: list | None = None) -> list:
    """MXAsymQuantizer.forward for several 2-D weights of ONE width and dtype in a single launch (the
    row-resident kernel; mxq_fakequant_fwd_multi).  pooled_mask: uint8[cols / group] from
    allocate_group_bits -- which groups share the row's 4-bit statistic (None = the positional recipe)."""
    import ctypes as C
    if not xs:
        return []
    L.require_cuda(*xs, pooled_mask)
    cols, dt = xs[0].shape[1], xs[0].dtype
    if any(x.dim() != 2 or x.shape[1] != cols or x.dtype != dt for x in xs):
        raise ValueError("fakequant_fwd_multi: 2-D tensors of one width and dtype")
    xs = [x.contiguous() for x in xs]
    if outs is None:
        outs = [torch.empty_like(x) for x in xs]
    n = len(xs)
    xa = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
    oa = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    ra = (C.c_int64 * n)(*[x.shape[0] for x in xs])
    with L.on(xs[0]) as st:
        rc = L.lib().mxq_fakequant_fwd_multi(xa, oa, ra, n, cols, L.dtype_enum(xs[0]), group, num_bits,
                                             L.ptr(pooled_mask), st)
    L.check(rc, "mxq_fakequant_fwd_multi")
    return outs


def ste_bwd(grad_out: torch.Tensor, x: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """MXAsymQuantizer.backward (utils_quant.py:464-475)."""
    L.require_cuda(grad_out, x)
    if grad_out.dtype != x.dtype or grad_out.shape != x.shape:
        raise ValueError("grad_out and x must match in dtype and shape")
    g = grad_out.contiguous()
    xc = x.contiguous()
    gi = torch.empty_like(g)
    with L.on(grad_out) as st:
        rc = L.lib().mxq_ste_bwd(L.ptr(g), L.ptr(xc), L.ptr(gi), g.numel(), L.dtype_enum(g),
                                 float(lo), float(hi), st)
    L.check(rc, "mxq_ste_bwd")
    return gi


def segquant_plan(shape, mode: str, layerwise: bool):
    """How Sym/AsymQuantizer (utils_quant.py:31-199) segment a tensor of `shape`:
    returns (nseg, seglen, period, valid)."""
    G = 128 if mode == "sym" else 8                       # :58 / :146
    n = 1
    for d in shape:
        n *= int(d)
    if layerwise:
        return 1, n, 1, 1                                  # :50-51 / :130-132
    nd = len(shape)
    if nd == 2:
        N, K = shape
        if K % G:
            # trailing K % G columns keep zero statistics in the reference (:59-68); rare and
            # not expressible as equal segments
            raise NotImplementedError(f"in_features {K} must be a multiple of the group size {G}")
        return N * (K // G), G, 1, 1
    if nd == 3:
        # the loop slices DIM 1 (tokens) with column-group indices and reduces over dim -1:
        # one statistic per token over all C channels, only for tokens < (C // G) * G
        B, T, Cc = shape
        live = min(T, (Cc // G) * G)
        return B * T, Cc, (T if live < T else 1), (live if live < T else 1)
    if nd == 4:
        B, H, T, D = shape
        return B * H, T * D, 1, 1                          # :72-79 / :171-187
    raise ValueError("Sym/AsymQuantizer: unsupported tensor rank")   # reference raises ValueError


def segquant_fwd(x: torch.Tensor, mode: str, bits: int, nseg: int, seglen: int, period: int = 1,
                 valid: int = 1, workspace=None) -> torch.Tensor:
    """Segment-statistic fake quantization (SymQuantizer / AsymQuantizer forward)."""
    L.require_cuda(x)
    x = x.contiguous()
    if nseg * seglen != x.numel():
        raise ValueError("nseg * seglen must equal the number of elements")
    out = torch.empty_like(x)
    need = L.lib().mxq_segquant_workspace_bytes(nseg, seglen, L.dtype_enum(x))
    if workspace is None or workspace.numel() < need:
        workspace = _ws(need, x.device)
    with L.on(x) as st:
        rc = L.lib().mxq_segquant_fwd(L.ptr(x), L.ptr(out), nseg, seglen, L.dtype_enum(x),
                                      0 if mode == "sym" else 1, int(bits), int(period), int(valid),
                                      L.ptr(workspace), workspace.numel(), st)
    L.check(rc, "mxq_segquant_fwd")
    return out


def colsumsq(X: torch.Tensor, out: torch.Tensor | None = None, prev_scale: float = 0.0,
             add_scale: float = 1.0, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """out = prev_scale*out + add_scale * sum_tokens X^2 (per column, fp32)."""
    L.require_cuda(X, out)
    X2 = X.reshape(-1, X.shape[-1]).contiguous()
    tokens, cols = X2.shape
    accumulate = out is not None
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=X.device)
    need = L.lib().mxq_colsumsq_workspace_bytes(tokens, cols)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = _ws(need, X.device)
    with L.on(X) as st:
        rc = L.lib().mxq_colsumsq(L.ptr(X2), tokens, cols, L.dtype_enum(X2), L.ptr(out),
                                  float(prev_scale), float(add_scale), int(accumulate),
                                  L.ptr(workspace), workspace.numel() * workspace.element_size(),
                                  st)
    L.check(rc, "mxq_colsumsq")
    return out


def wanda_metric(W: torch.Tensor, scaler_row: torch.Tensor) -> torch.Tensor:
    """|W| * sqrt(scaler_row) (mxq_quant/lib/prune.py:177)."""
    L.require_cuda(W, scaler_row)
    W = W.contiguous()
    out = torch.empty(W.shape, dtype=torch.float32, device=W.device)
    with L.on(W) as st:
        rc = L.lib().mxq_wanda_metric(L.ptr(W), L.ptr(scaler_row.contiguous().float()), L.ptr(out),
                                      W.shape[0], W.shape[1], L.dtype_enum(W), st)
    L.check(rc, "mxq_wanda_metric")
    return out


def allocate_group_bits(W: torch.Tensor, scaler_row: torch.Tensor | None = None, group: int = 16,
                        low_bits: int = 2, return_importance: bool = False):
    """Importance-driven 2/4-bit allocation: of every 4 consecutive `group`-column groups the one with
    the largest sum of |W| * sqrt(scaler_row) (Wanda metric, prune.py:177) becomes the pooled 4-bit
    group.  Returns the uint8 group_bits mask for fakequant_fwd / ptq_quant (and the fp64 importances)."""
    L.require_cuda(W, scaler_row)
    if W.dtype != torch.float16:
        raise TypeError("allocate_group_bits expects fp16 weights")
    W = W.contiguous()
    rows, cols = W.shape
    if cols % (4 * group):
        raise ValueError(f"cols={cols} must be a multiple of 4*group={4 * group}")
    sr = None if scaler_row is None else scaler_row.contiguous().float()
    gb = torch.empty(cols // group, dtype=torch.uint8, device=W.device)
    imp = torch.empty(cols // group, dtype=torch.float64, device=W.device) if return_importance else None
    ws = _ws(L.lib().mxq_allocate_bits_workspace_bytes(cols), W.device)
    with L.on(W) as st:
        rc = L.lib().mxq_allocate_bits(L.ptr(W), L.ptr(sr), rows, cols, group, low_bits, L.ptr(gb), L.ptr(imp),
                                       L.ptr(ws), ws.numel(), st)
    L.check(rc, "mxq_allocate_bits")
    return (gb, imp) if return_importance else gb


def importance_permutation(group_bits: torch.Tensor) -> torch.Tensor:
    """Group permutation that moves the pooled (4-bit) group of every 4 consecutive groups -- as
    chosen by `allocate_group_bits` -- to the last slot, where the positional packed layout keeps its
    4-bit columns; the other three keep their order.  Returns int32[ngroups] with
    packed group g = original group perm[g] (host-side integer bookkeeping, on the mask's device)."""
    gb = group_bits.detach().to("cpu").view(-1, 4)
    pooled = (gb & POOL) != 0
    if not bool((pooled.sum(dim=1) == 1).all()):
        raise ValueError("exactly one pooled group per 4 consecutive groups is required")
    j = pooled.int().argmax(dim=1)                                  # pooled slot of every block
    slots = torch.arange(4).expand(gb.shape[0], 4)
    rest = slots[slots != j[:, None]].view(-1, 3)
    order = torch.cat([rest, j[:, None]], dim=1)                     # [nblk, 4]
    perm = (order + 4 * torch.arange(gb.shape[0])[:, None]).reshape(-1).to(torch.int32)
    return perm.to(group_bits.device)


def gather_groups(t: torch.Tensor, group_perm: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[..., 16 g : 16 g + 16] = t[..., 16 perm[g] : ...] for fp16 rows (mxq_gather_groups)."""
    L.require_cuda(t, group_perm, out)
    if t.dtype != torch.float16 or t.shape[-1] % 16 or group_perm.dtype != torch.int32 or \
            group_perm.numel() != t.shape[-1] // 16:
        raise ValueError("gather_groups: fp16 [..., cols] with cols % 16 == 0 and int32 perm[cols / 16]")
    t2 = t.reshape(-1, t.shape[-1]).contiguous()
    if out is None:
        out = torch.empty_like(t2)
    with L.on(t2) as st:
        rc = L.lib().mxq_gather_groups(L.ptr(t2), L.ptr(group_perm.contiguous()), L.ptr(out), t2.shape[0], t2.shape[1], st)
    L.check(rc, "mxq_gather_groups")
    return out.reshape(t.shape)


def ptq_quant(W: torch.Tensor, colstat: torch.Tensor | None = None, low_bits: int = 2,
              group: int = 16, group_bits=None, return_codes: bool = False, out=None,
              workspace=None):
    """MXQGPT.fasterquant (mxq_quant/lib/mxqgpt.py:387-448): fp16 in, fp16 out.  `group` = width of
    the 2-bit groups: 16 (blocksize=16, prune.py:409), or 32 / 48 for the reference recipe
    (blocksize 32 / >= 48: groups inside the 48 low columns of every 64-column block)."""
    L.require_cuda(W, colstat, group_bits)
    if W.dtype != torch.float16:
        raise TypeError("ptq_quant expects fp16 weights (the reference casts back to the weight dtype)")
    W = W.contiguous()
    rows, cols = W.shape
    Wq = torch.empty_like(W) if out is None else out
    codes = torch.empty(W.shape, dtype=torch.uint8, device=W.device) if return_codes else None
    need = L.lib().mxq_ptq_workspace_bytes(rows, cols)
    if workspace is None or workspace.numel() < need:
        workspace = _ws(need, W.device)
    with L.on(W) as st:
        rc = L.lib().mxq_ptq_quant(L.ptr(W), L.ptr(Wq), L.ptr(codes), L.ptr(colstat), rows, cols, group,
                                   low_bits, L.ptr(group_bits), L.ptr(workspace), workspace.numel(),
                                   st)
    L.check(rc, "mxq_ptq_quant")
    return (Wq, codes) if return_codes else Wq


def rowquant(x: torch.Tensor, bits: int, qq_scale_bits: int | None = 4):
    """Quantizer.find_params + quantize_dequantize / quantize (quantizer.py:61-121,149-173) on an
    fp32 [rows, cols] matrix.  Returns (y, codes, scale, zero)."""
    L.require_cuda(x)
    x = x.contiguous().float()
    rows, cols = x.shape
    y = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    scale = torch.empty(rows, dtype=torch.float32, device=x.device)
    zero = torch.empty(rows, dtype=torch.float32, device=x.device)
    with L.on(x) as st:
        rc = L.lib().mxq_rowquant(L.ptr(x), L.ptr(y), L.ptr(codes), L.ptr(scale), L.ptr(zero), rows,
                                  cols, bits, qq_scale_bits or 0, st)
    L.check(rc, "mxq_rowquant")
    return y, codes, scale, zero


def packed_shapes(OC: int, IC: int) -> dict:
    if IC % 64 or OC % 8:
        raise ValueError("IC must be a multiple of 64 and OC a multiple of 8")
    nblk = IC // 64
    nch = (nblk + 63) // 64
    return dict(weight=((OC, nblk * 4), torch.int32), weight_last=((OC, nblk), torch.int32),
                zeros_and_scales=((OC, 32 * nch), torch.int32),
                zeros_2nd=((OC // 4, 32 * nch), torch.int32),
                scales_2nd=((OC // 4, 3 * nblk), torch.float16),
                scales_4b=((OC,), torch.float16), zeros_4b=((OC // 8,), torch.int32))


def alloc_packed(OC: int, IC: int, device) -> dict:
    return {k: torch.empty(s, dtype=d, device=device) for k, (s, d) in packed_shapes(OC, IC).items()}


def pack(W: torch.Tensor, colstat: torch.Tensor | None = None, out: dict | None = None,
         workspace=None, group_perm: torch.Tensor | None = None) -> dict:
    """fp16 W[OC, IC] -> packed mixed 2/4-bit tensors (layout of gemv_mxq_cuda.cu:39-208).
    group_perm (int32[IC/16], e.g. from importance_permutation): the column order the weights are
    packed in; the consumer must be given the same permutation (gemv(..., group_perm=...), MXQLinear)."""
    L.require_cuda(W, colstat, group_perm)
    if W.dtype != torch.float16:
        raise TypeError("pack expects fp16 weights")
    W = W.contiguous()
    if group_perm is not None:
        W = gather_groups(W, group_perm)
        if colstat is not None:
            colstat = colstat.float().view(-1, 16)[group_perm.long()].reshape(-1).contiguous()
    OC, IC = W.shape
    if out is None:
        out = alloc_packed(OC, IC, W.device)
    need = L.lib().mxq_pack_workspace_bytes(OC, IC)
    if workspace is None or workspace.numel() < need:
        workspace = _ws(need, W.device)
    with L.on(W) as st:
        rc = L.lib().mxq_pack(L.ptr(W), L.ptr(colstat), OC, IC, L.packed_struct(out), L.ptr(workspace),
                              workspace.numel(), st)
    L.check(rc, "mxq_pack")
    return out


def ptq_quant_pack(W: torch.Tensor, colstat: torch.Tensor | None = None, out=None, packed=None,
                   workspace=None, return_codes: bool = False):
    """fasterquant + pack in one pass over W: returns (Wq fp16, packed dict[, codes])."""
    L.require_cuda(W, colstat)
    if W.dtype != torch.float16:
        raise TypeError("ptq_quant_pack expects fp16 weights")
    W = W.contiguous()
    rows, cols = W.shape
    Wq = torch.empty_like(W) if out is None else out
    if packed is None:
        packed = alloc_packed(rows, cols, W.device)
    codes = torch.empty(W.shape, dtype=torch.uint8, device=W.device) if return_codes else None
    need = L.lib().mxq_ptq_workspace_bytes(rows, cols)
    if workspace is None or workspace.numel() < need:
        workspace = _ws(need, W.device)
    with L.on(W) as st:
        rc = L.lib().mxq_ptq_quant_pack(L.ptr(W), L.ptr(Wq), L.ptr(codes), L.ptr(colstat), rows, cols,
                                        L.packed_struct(packed), L.ptr(workspace), workspace.numel(),
                                        st)
    L.check(rc, "mxq_ptq_quant_pack")
    return (Wq, packed, codes) if return_codes else (Wq, packed)


def _packed_dims(p: dict):
    OC = p["weight"].shape[0]
    IC = p["weight"].shape[1] * 16
    return OC, IC


def _check_packed(p: dict):
    OC, IC = _packed_dims(p)
    for k, (shape, dt) in packed_shapes(OC, IC).items():
        t = p[k]
        if tuple(t.shape) != tuple(shape) or t.dtype != dt or not t.is_contiguous():
            raise ValueError(f"packed tensor '{k}' must be contiguous {dt} {tuple(shape)}, got "
                             f"{t.dtype} {tuple(t.shape)}")
    L.require_cuda(*p.values())
    return OC, IC


def unpack(p: dict, dtype=torch.float32) -> torch.Tensor:
    OC, IC = _check_packed(p)
    out = torch.empty((OC, IC), dtype=dtype, device=p["weight"].device)
    with L.on(out) as st:
        rc = L.lib().mxq_unpack(L.packed_struct(p), OC, IC, L.ptr(out), L.dtype_enum(out), st)
    L.check(rc, "mxq_unpack")
    return out


def gemv(x: torch.Tensor, p: dict, out: torch.Tensor | None = None, validate: bool = True,
         pdl: bool = False, group_perm: torch.Tensor | None = None):
    """y[B, OC] = x[B, IC] @ dequant(W)^T, fp16.  pdl=True: programmatic dependent launch -- the
    kernel copies its packed weights into shared memory while the previous kernel of the stream is
    still running, which is only correct if that kernel does not write them (a decode chain over
    resident weights); x and y are touched after the dependency wait in either mode."""
    OC, IC = _check_packed(p) if validate else _packed_dims(p)
    L.require_cuda(x, p["weight"], out)
    if x.dtype != torch.float16 or x.dim() != 2 or x.shape[1] != IC:
        raise ValueError(f"x must be fp16 [B, {IC}]")
    x = x.contiguous()
    B = x.shape[0]
    if out is None:
        out = torch.empty((B, OC), dtype=torch.float16, device=x.device)
    if group_perm is not None:
        import ctypes as C
        if group_perm.dtype != torch.int32 or group_perm.numel() != IC // 16 or group_perm.device != x.device:
            raise ValueError(f"group_perm must be int32[{IC // 16}] on the activations' device")
        warr = (L.PackedC * 1)(L.packed_struct(p))
        yarr = (C.c_void_p * 1)(out.data_ptr())
        with L.on(x) as st:
            rc = L.lib().mxq_gemv_grouped_perm(L.ptr(x), warr, yarr, 1, B, IC, OC, L.ptr(group_perm.contiguous()),
                                               0 if pdl else 1, st)
        L.check(rc, "mxq_gemv_grouped_perm")
        return out
    with L.on(x) as st:
        rc = L.lib().mxq_gemv_ex(L.ptr(x), L.packed_struct(p), L.ptr(out), B, IC, OC,
                                 0 if pdl else 1, st)
    L.check(rc, "mxq_gemv")
    return out


def gemv_grouped(x: torch.Tensor, ps: list, outs: list | None = None, pdl: bool = False, validate: bool = True):
    """Decode GEMV of up to 4 packed linears of one shape sharing the activation x (q/k/v, gate/up)
    in ONE launch.  Returns the list of fp16 [B, OC] outputs."""
    import ctypes as C
    if not 1 <= len(ps) <= 4:
        raise ValueError("gemv_grouped takes 1..4 packed linears")
    dims = [(_check_packed(p) if validate else _packed_dims(p)) for p in ps]
    if len(set(dims)) != 1:
        raise ValueError("all linears of a group must have the same [OC, IC]")
    OC, IC = dims[0]
    L.require_cuda(x, *[q["weight"] for q in ps])
    if x.dtype != torch.float16 or x.dim() != 2 or x.shape[1] != IC:
        raise ValueError(f"x must be fp16 [B, {IC}]")
    x = x.contiguous()
    B = x.shape[0]
    if outs is None:
        outs = [torch.empty((B, OC), dtype=torch.float16, device=x.device) for _ in ps]
    warr = (L.PackedC * len(ps))(*[L.packed_struct(p) for p in ps])
    yarr = (C.c_void_p * len(ps))(*[o.data_ptr() for o in outs])
    with L.on(x) as st:
        rc = L.lib().mxq_gemv_grouped(L.ptr(x), warr, yarr, len(ps), B, IC, OC, 0 if pdl else 1, st)
    L.check(rc, "mxq_gemv_grouped")
    return outs


class GemvChain:
    """A sequence of batch-1 decode GEMVs as ONE persistent launch (mxq_gemv_chain_*).

    jobs: list of (x, packed, y, dep) -- x fp16 [IC] or [1, IC], y fp16 [OC] or [1, OC] (buffers owned by the
    caller and kept alive by this object), dep = -1 (x ready at launch) or the index of an earlier job
    whose y this job's x is (or depends on).  Jobs are otherwise unordered.  The plan is built once;
    `run()` launches it on the current stream (capturable in a CUDA graph).  Longer chains than
    MXQ_GEMV_CHAIN_MAX_JOBS are cut into several launches (stream order covers dependencies across them).
    Raises RuntimeError with MXQ_E_UNSUPPORTED for shapes the chain kernel does not take (IC % 256 or OC % 32):
    callers fall back to `gemv` per linear."""

    def __init__(self, jobs: list, validate: bool = True):
        import ctypes as C
        if not jobs:
            raise ValueError("GemvChain needs at least one job")
        lib = L.lib()
        self._keep = []
        self._launches = []
        dev = jobs[0][0].device
        cap = L.GEMV_CHAIN_MAX_JOBS
        for lo in range(0, len(jobs), cap):
            part = jobs[lo:lo + cap]
            arr = (L.GemvJobC * len(part))()
            for i, (x, p, y, dep) in enumerate(part):
                OC, IC = _check_packed(p) if validate else _packed_dims(p)
                L.require_cuda(x, y, p["weight"])
                if x.device != dev:
                    raise RuntimeError("all jobs of a chain must live on one device")
                if x.dtype != torch.float16 or x.numel() != IC or not x.is_contiguous():
                    raise ValueError(f"job {lo + i}: x must be contiguous fp16 with {IC} elements")
                if y.dtype != torch.float16 or y.numel() != OC or not y.is_contiguous():
                    raise ValueError(f"job {lo + i}: y must be contiguous fp16 with {OC} elements")
                d = -1 if dep is None else int(dep)
                if d >= lo + i:
                    raise ValueError(f"job {lo + i}: dep must name an earlier job")
                # a dependency on a job of an earlier launch is covered by stream order
                arr[i] = L.GemvJobC(x.data_ptr(), y.data_ptr(), L.packed_struct(p), IC, OC, d - lo if d >= lo else -1, 0)
                self._keep.append((x, p, y))
            host = torch.zeros(int(lib.mxq_gemv_chain_plan_bytes()) + 64, dtype=torch.uint8)
            host = host[(-host.data_ptr()) % 64:][:int(lib.mxq_gemv_chain_plan_bytes())]      # 64-byte aligned view
            with L.on(x):                         # the tensor maps are encoded against the jobs' device
                L.check(lib.mxq_gemv_chain_plan(arr, len(part), host.data_ptr()), "mxq_gemv_chain_plan")
            sync = torch.zeros(L.GEMV_CHAIN_SYNC_WORDS, dtype=torch.int32, device=dev)
            self._launches.append((host, host.to(dev), sync))
        self.device = dev
        self.n = len(jobs)

    def run(self, pdl: bool = False):
        """pdl=True (MXQ_GEMV_CHAIN_PDL, chains without dependencies): the launch overlaps the tail of the
        previous kernel of the stream with its own set-up and weight prefetch; only for callers whose previous
        kernel does not write this chain's packed weights (a decode loop: the previous step's chain)."""
        lib = L.lib()
        with L.on(self._launches[0][2]) as st:
            for host, plan_dev, sync in self._launches:
                rc = lib.mxq_gemv_chain_run(host.data_ptr(), plan_dev.data_ptr(), sync.data_ptr(), L.GEMV_CHAIN_PDL if pdl else 0, st)
                L.check(rc, "mxq_gemv_chain_run")


def gemm_workspace_bytes(M: int, IC: int, OC: int) -> int:
    return max(int(L.lib().mxq_gemm_workspace_bytes(M, IC, OC)), 16)


def gemm_workspace(M: int, IC: int, OC: int, device) -> torch.Tensor:
    """Workspace for `gemm` at this shape: the fp32 exchange buffer of the K-split tail tiles
    (mxq_gemm_workspace_bytes; need not be initialised)."""
    return _ws(gemm_workspace_bytes(M, IC, OC), device)


def gemm(x: torch.Tensor, p: dict, out: torch.Tensor | None = None, workspace=None,
         validate: bool = True, split_k: bool = True):
    """Prefill: y[M, OC] = x[M, IC] @ dequant(W)^T on tcgen05/TMEM, fp16 in/out, fp32 accumulate.
    split_k=False passes no workspace: whole tiles only, i.e. every output element is one
    K-ordered fp32 accumulation (bit-identical to the sharded / fused-exchange paths)."""
    OC, IC = _check_packed(p) if validate else _packed_dims(p)
    L.require_cuda(x, p["weight"], out)
    if x.dtype != torch.float16 or x.dim() != 2 or x.shape[1] != IC:
        raise ValueError(f"x must be fp16 [M, {IC}]")
    x = x.contiguous()
    M = x.shape[0]
    if out is None:
        out = torch.empty((M, OC), dtype=torch.float16, device=x.device)
    if split_k:
        need = L.lib().mxq_gemm_workspace_bytes(M, IC, OC)
        if workspace is None or workspace.numel() < need:
            workspace = _ws(need, x.device)
        wptr, wbytes = L.ptr(workspace), workspace.numel()
    else:
        wptr, wbytes = None, 0
    with L.on(x) as st:
        rc = L.lib().mxq_gemm(L.ptr(x), L.packed_struct(p), L.ptr(out), M, IC, OC, wptr, wbytes, st)
    L.check(rc, "mxq_gemm")
    return out


def gemm_scatter(x: torch.Tensor, p: dict, peer_ptrs: list, ldy: int, col0: int, workspace=None):
    """Column-sharded GEMM whose epilogue stores the [M, OC_local] tile at column `col0` of every
    [M, ldy] fp16 buffer in `peer_ptrs` (device addresses; peers mapped over NVLink)."""
    import ctypes as C
    OC, IC = _packed_dims(p)
    L.require_cuda(x, p["weight"])
    if x.dtype != torch.float16 or x.dim() != 2 or x.shape[1] != IC:
        raise ValueError(f"x must be fp16 [M, {IC}]")
    x = x.contiguous()
    arr = (C.c_void_p * len(peer_ptrs))(*[int(a) for a in peer_ptrs])
    with L.on(x) as st:
        rc = L.lib().mxq_gemm_scatter(L.ptr(x), L.packed_struct(p), arr, len(peer_ptrs), x.shape[0], IC, OC,
                                      ldy, col0, L.ptr(workspace), 0 if workspace is None else workspace.numel(),
                                      st)
    L.check(rc, "mxq_gemm_scatter")


def gemm_multicast(x: torch.Tensor, p: dict, multicast_ptr: int, ldy: int, col0: int, workspace=None):
    """Column-sharded GEMM whose epilogue stores the [M, OC_local] tile at column `col0` of the
    symmetric [M, ldy] fp16 buffer behind the NVSwitch multicast address `multicast_ptr`: one
    multimem.st per 16 bytes lands in every rank's copy."""
    OC, IC = _packed_dims(p)
    L.require_cuda(x, p["weight"])
    if x.dtype != torch.float16 or x.dim() != 2 or x.shape[1] != IC:
        raise ValueError(f"x must be fp16 [M, {IC}]")
    if not multicast_ptr:
        raise RuntimeError("gemm_multicast needs a multicast mapping (symmetric memory without NVSwitch multicast support)")
    x = x.contiguous()
    with L.on(x) as st:
        rc = L.lib().mxq_gemm_multicast(L.ptr(x), L.packed_struct(p), int(multicast_ptr), x.shape[0], IC, OC,
                                        ldy, col0, L.ptr(workspace), 0 if workspace is None else workspace.numel(),
                                        st)
    L.check(rc, "mxq_gemm_multicast")


def gemm_dense(x: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """Diagnostic: y = x @ W^T with dense fp16 W through the same tcgen05 pipeline as `gemm`."""
    L.require_cuda(x, W)
    x, W = x.contiguous(), W.contiguous()
    M, IC = x.shape
    OC = W.shape[0]
    out = torch.empty((M, OC), dtype=torch.float16, device=x.device)
    with L.on(x) as st:
        rc = L.lib().mxq_gemm_dense(L.ptr(x), L.ptr(W), L.ptr(out), M, IC, OC, st)
    L.check(rc, "mxq_gemm_dense")
    return out


def awq_gemv(x: torch.Tensor, kernel: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor,
             group_size: int) -> torch.Tensor:
    """AWQ uniform 4-bit GEMV (gemv_cuda.cu:346-399)."""
    L.require_cuda(x, kernel, scales, zeros)
    x = x.contiguous()
    B, IC = x.shape
    OC = kernel.shape[0]
    out = torch.empty((B, OC), dtype=torch.float16, device=x.device)
    with L.on(x) as st:
        rc = L.lib().mxq_awq_gemv(L.ptr(x), L.ptr(kernel.contiguous()), L.ptr(scales.contiguous()),
                                  L.ptr(zeros.contiguous()), L.ptr(out), B, IC, OC, group_size,
                                  st)
    L.check(rc, "mxq_awq_gemv")
    return out


def awq_gemm(x: torch.Tensor, kernel: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor,
             workspace: torch.Tensor | None = None) -> torch.Tensor:
    """AWQ uniform 4-bit prefill GEMM (gemm_cuda_gen.cu:424-478): x fp16 [M, IC], kernel int32
    [IC, OC/8], scales fp16 [IC/G, OC], zeros int32 [IC/G, OC/8] -> fp16 [M, OC]."""
    L.require_cuda(x, kernel, scales, zeros)
    if x.dtype != torch.float16 or x.dim() != 2 or kernel.dim() != 2 or kernel.shape[0] != x.shape[1]:
        raise ValueError("x must be fp16 [M, IC] and kernel int32 [IC, OC/8]")
    x = x.contiguous()
    M, IC = x.shape
    OC = kernel.shape[1] * 8
    if scales.dim() != 2 or scales.shape[1] != OC or IC % scales.shape[0]:
        raise ValueError("scales must be fp16 [IC/G, OC]")
    G = IC // scales.shape[0]
    out = torch.empty((M, OC), dtype=torch.float16, device=x.device)
    need = L.lib().mxq_awq_gemm_workspace_bytes(IC, OC)
    if workspace is None or workspace.numel() < need:
        workspace = _ws(need, x.device)
    with L.on(x) as st:
        rc = L.lib().mxq_awq_gemm(L.ptr(x), L.ptr(kernel.contiguous()), L.ptr(scales.contiguous()),
                                  L.ptr(zeros.contiguous()), L.ptr(out), M, IC, OC, G, L.ptr(workspace),
                                  workspace.numel(), st)
    L.check(rc, "mxq_awq_gemm")
    return out
