"""CPU oracle for the MXQ quantization hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithms.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``mxq_b200/`` imports it, and the product path raises when the
CUDA library is missing (there is no CPU fallback).

Parity status (see DESIGN.md "Oracle"):
  * fakequant_fwd / ste_bwd / fasterquant / scaler_row / wanda_metric are PINNED:
    ``oracle/gen_golden.py`` runs the unmodified reference Python
    (/root/reference, importable in the build container only) on seeded inputs
    and stores input/output vectors under ``tests/golden/``;
    ``tests/test_oracle_golden.py`` checks this file against them bit for bit.
  * decode_mxq / gemv_mxq are PINNED to the reference's only known-answer test
    (cuda_kernel/test_correct_gemv.py:19-53) and, on the GPU box, to the
    reference CUDA kernel compiled from its own source (oracle/_ref).
  * pack_mxq (weights -> packed layout) is "parity unpinned": the reference has
    no producer for that layout.  The encode policy is
    ours; the decode formula is the contract.

All arithmetic below is IEEE fp32 done op by op (numpy float32), matching the
reference's eager ATen kernels: every intermediate is rounded to the tensor
dtype (fp32 / bf16 / fp16), ``round`` is round-half-to-even.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# dtype emulation
# --------------------------------------------------------------------------------------

def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (nearest even) and return the value as fp32."""
    x = np.ascontiguousarray(x, dtype=F32)
    bits = x.view(np.uint32)
    nan = np.isnan(x)
    bias = np.uint32(0x7FFF) + ((bits >> np.uint32(16)) & np.uint32(1))
    out = ((bits + bias) & np.uint32(0xFFFF0000)).view(F32)
    if nan.any():
        out = np.where(nan, x, out)
    return out


def round_fp16(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=F32).astype(np.float16).astype(F32)


def rounder(dtype: str):
    """dtype in {'fp32','bf16','fp16'} -> function rounding an fp32 array to that dtype."""
    if dtype == "fp32":
        return lambda a: np.asarray(a, dtype=F32)
    if dtype == "bf16":
        return round_bf16
    if dtype == "fp16":
        return round_fp16
    raise ValueError(dtype)


# --------------------------------------------------------------------------------------
# bit-width recipe
# --------------------------------------------------------------------------------------

POOL = 0x80   # group_bits flag: this group belongs to the per-row pool (shares one min/max per row)


def reference_group_bits(cols: int, group: int = 16, low_bits: int = 2) -> np.ndarray:
    """The reference's positional recipe: of every 4 consecutive groups the first three are
    low-bit with their own min/max and the last is 4-bit and pooled per row
    (LLM-QAT/models/utils_quant.py:340-385 with groupsize=16, ratio_2b=6/8, num_4b=16;
    mxq_quant/lib/mxqgpt.py:404-419).  Encoding: low 7 bits = bit-width, bit 7 = pooled."""
    if cols % (4 * group):
        raise ValueError(f"cols={cols} must be a multiple of 4*group={4 * group}")
    ng = cols // group
    gb = np.full(ng, low_bits, dtype=np.uint8)
    gb[3::4] = POOL | 4
    return gb


def _split_bits(group_bits):
    group_bits = np.asarray(group_bits, dtype=np.uint8)
    pool = (group_bits & POOL) != 0
    bits = (group_bits & 0x7F).astype(np.int64)
    if pool.any() and len(set(bits[pool].tolist())) != 1:
        raise ValueError("pooled groups must share one bit-width")
    return bits, pool


# --------------------------------------------------------------------------------------
# (a-1) MXAsymQuantizer.forward   LLM-QAT/models/utils_quant.py:315-462
# --------------------------------------------------------------------------------------

def fakequant_fwd(x, dtype: str = "fp32", num_bits: int = 2, group: int = 16,
                  group_bits: np.ndarray | None = None, return_aux: bool = False):
    """Restates utils_quant.py:334-385 (per-group alpha/beta/s) and :456-460 (quant/dequant).

    x: [N, K] array holding values representable in `dtype` (stored as fp32).
    Low-bit groups: alpha = max-min, beta = min over the `group` columns of one row
    (:357-365), s = 2**num_bits-1 (:366).  4-bit groups: all such columns of a row are pooled
    (:347,368) and share one alpha/beta computed in fp32 (:369-377; W_4b is an fp32 tensor)
    then stored into the dtype tensor (:383-384), s = 15 (:385).
    out = round(((x-beta)/(alpha+1e-8))*s)/s*(alpha+1e-8)+beta, each op rounded to dtype.
    """
    rnd = rounder(dtype)
    x = np.asarray(x, dtype=F32)
    N, K = x.shape
    if K % group:
        raise ValueError("cols must be a multiple of group")
    ng = K // group
    if group_bits is None:
        group_bits = reference_group_bits(K, group, num_bits)
    bits, pool = _split_bits(group_bits)
    assert bits.shape == (ng,)
    xg = x.reshape(N, ng, group)
    gmin = xg.min(axis=2)
    gmax = xg.max(axis=2)
    alpha = rnd(gmax - gmin)                       # dtype subtract (:358-361)
    beta = gmin.copy()
    s = np.broadcast_to((2.0 ** bits - 1).astype(F32)[None, :], (N, ng)).copy()
    if pool.any() and N > 0:
        pmin = gmin[:, pool].min(axis=1, keepdims=True)
        pmax = gmax[:, pool].max(axis=1, keepdims=True)
        a4 = rnd((pmax - pmin).astype(F32))        # fp32 subtract then cast on assignment
        alpha[:, pool] = a4
        beta[:, pool] = pmin
    # the CPU kernel casts the python scalar to the tensor dtype before adding (bf16: 1.0012e-8)
    a = rnd(alpha + rnd(F32(1e-8)))                # alpha + 1e-8 (:456)
    a3, b3, s3 = a[:, :, None], beta[:, :, None], s[:, :, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = rnd(xg - b3)
        t = rnd(t / a3)                            # input_normalized (:456)
        t = rnd(t * s3)
        q = np.rint(t).astype(F32)                 # torch.round = half-to-even (:458)
        t = rnd(q / s3)
        t = rnd(t * a3)
        out = rnd(t + b3)                          # (:460)
    out = out.reshape(N, K)
    if return_aux:
        return out, q.reshape(N, K), alpha, beta, s
    return out


def allocate_group_bits(W16, scaler_row=None, group: int = 16, low_bits: int = 2):
    """Importance-driven allocation: importance[g] = sum_{r, c in g}
    |W[r,c]| * sqrt(scaler_row[c]) -- the Wanda metric of mxq_quant/lib/prune.py:177 summed per
    column group -- and of every 4 consecutive groups the most important one becomes the pooled
    4-bit group (ties: lowest index); the reference itself always takes the last one
    (utils_quant.py:340-385).  Exact arithmetic: |fp16| column sums as integers in units of 2^-24,
    then fp64 in column order.  Returns (group_bits uint8[K/group], importance float64[K/group])."""
    W = np.asarray(W16, dtype=np.float16)
    N, K = W.shape
    if K % (4 * group):
        raise ValueError("cols must be a multiple of 4*group")
    mag = np.abs(W.astype(np.float64)) * 16777216.0
    mag = np.where(np.isfinite(mag), mag, 0.0)
    colabs = mag.astype(np.int64).sum(axis=0)
    a = colabs.astype(np.float64) * (1.0 / 16777216.0)
    w = np.ones(K) if scaler_row is None else np.sqrt(np.asarray(scaler_row, dtype=np.float32).astype(np.float64))
    term = (a * w).reshape(K // group, group)
    imp = np.zeros(K // group)
    for j in range(group):                      # fixed order, like the kernel
        imp = imp + term[:, j]
    blocks = imp.reshape(-1, 4)
    arg = np.argmax(blocks, axis=1)             # first maximum
    gb = np.full((K // group // 4, 4), low_bits, dtype=np.uint8)
    gb[np.arange(gb.shape[0]), arg] = POOL | 4
    return gb.reshape(-1), imp


# --------------------------------------------------------------------------------------
# (f-1) SymQuantizer / AsymQuantizer.forward   LLM-QAT/models/utils_quant.py:31-95, 98-199
# --------------------------------------------------------------------------------------

def _act_segments(shape, group: int, layerwise: bool):
    """The statistic layout of utils_quant.py:50-81 / :130-187 as (view shape with the reduced
    axis last, live mask over the leading axes or None).
      layerwise            : one statistic for the tensor (:50-51, :130-132)
      2-D [N, K]           : the loop slices dim 1 = columns: per (row, group of `group` columns)
      3-D [B, T, C]        : the SAME loop slices dim 1 = TOKENS i*G:(i+1)*G for i < C // G and
                             reduces over dim -1: per-token statistic over all C channels, tokens
                             >= (C // G) * G keep the zero-initialised statistic (:56-64, :144-157)
      4-D [B, H, T, D]     : per (b, h) over T*D (:72-79, :171-187)"""
    shape = tuple(int(d) for d in shape)
    if layerwise:
        return (1, int(np.prod(shape))), None
    if len(shape) == 2:
        N, K = shape
        if K % group:
            raise NotImplementedError("trailing K % group columns keep zero statistics")
        return (N * (K // group), group), None
    if len(shape) == 3:
        B, T, Cc = shape
        live = (np.arange(T) < (Cc // group) * group)
        return (B * T, Cc), np.tile(live, B)
    if len(shape) == 4:
        B, H, T, D = shape
        return (B * H, T * D), None
    raise ValueError


def sym_quant(x, dtype: str = "fp32", num_bits: int = 8, layerwise: bool = False):
    """SymQuantizer.forward (utils_quant.py:38-87): max_input = max|x| per segment;
    s = (2**(bits-1) - 1) / (max_input + 1e-6) -- python-int / tensor dispatches to
    Tensor.__rdiv__ = tensor.reciprocal() * int, i.e. TWO rounded ops (:84);
    output = round(input * s).div(s + 1e-6) (:85)."""
    rnd = rounder(dtype)
    x = np.asarray(x, dtype=F32)
    view, live = _act_segments(x.shape, 128, layerwise)
    xs = x.reshape(view)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if xs.shape[0] and xs.shape[1]:
            m = np.abs(xs).max(axis=1, keepdims=True)
            m = np.where(np.isnan(xs).any(axis=1, keepdims=True), F32(np.nan), m)
        else:
            m = np.zeros((xs.shape[0], 1), dtype=F32)
        if live is not None:
            m = np.where(live[:, None], m, F32(0))
        qmax = rnd(F32(2 ** (num_bits - 1) - 1))   # python scalar, cast to the tensor dtype
        c6 = rnd(F32(1e-6))                        # python scalars are cast to the tensor dtype first
        d = rnd(m + c6)
        s = rnd(rnd(F32(1.0) / d) * qmax)
        t = rnd(xs * s)
        q = np.rint(t).astype(F32)
        out = rnd(q / rnd(s + c6))
    return out.reshape(x.shape)


def asym_quant(x, dtype: str = "fp32", num_bits: int = 8, layerwise: bool = False):
    """AsymQuantizer.forward (utils_quant.py:105-185): alpha = max - min, beta = min per
    segment; input_normalized = (input - beta) / (alpha + 1e-8) (:179); s = 2**bits - 1;
    quant_input = round(input_normalized * s).div(s) (:181); output = quant_input *
    (alpha + 1e-8) + beta (:183).  Every op rounded to the tensor dtype."""
    rnd = rounder(dtype)
    x = np.asarray(x, dtype=F32)
    view, live = _act_segments(x.shape, 8, layerwise)
    xs = x.reshape(view)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if xs.shape[0] and xs.shape[1]:
            mn = xs.min(axis=1, keepdims=True)
            mx = xs.max(axis=1, keepdims=True)
        else:
            mn = np.zeros((xs.shape[0], 1), dtype=F32)
            mx = mn.copy()
        if live is not None:
            mn = np.where(live[:, None], mn, F32(0))
            mx = np.where(live[:, None], mx, F32(0))
        alpha = rnd(mx - mn)
        a = rnd(alpha + rnd(F32(1e-8)))            # python scalars are cast to the tensor dtype first
        s = rnd(F32(2 ** num_bits - 1))            # python scalar, cast to the tensor dtype
        t = rnd(xs - mn)
        t = rnd(t / a)
        t = rnd(t * s)
        q = np.rint(t).astype(F32)
        t = rnd(q / s)
        t = rnd(t * a)
        out = rnd(t + mn)
    return out.reshape(x.shape)


# --------------------------------------------------------------------------------------
# (a-2) MXAsymQuantizer.backward   utils_quant.py:464-475
# --------------------------------------------------------------------------------------

def ste_bwd(grad_out, x, lo=-2.0, hi=2.0):
    """grad_in = grad_out.clone(); grad_in[x >= hi] = 0; grad_in[x <= lo] = 0 (:471-474).
    The comparison happens in the tensor dtype against the fp32 clip value (exact)."""
    g = np.array(grad_out, dtype=F32, copy=True)
    x = np.asarray(x, dtype=F32)
    g[x >= F32(hi)] = 0
    g[x <= F32(lo)] = 0
    return g


# --------------------------------------------------------------------------------------
# (a-8) WrappedGPT.add_batch + Wanda metric   layerwrapper.py:22-35, prune.py:177
# (a-5) MXQGPT.add_batch: only diag(H)==0 is consumed (mxqgpt.py:399-403); diag(H) is
#       2 * scaler_row up to summation order, and it is zero exactly when the column is zero.
# --------------------------------------------------------------------------------------

def colsumsq(X) -> np.ndarray:
    """Sum over tokens of x^2 per column, fp32 inputs promoted from fp16; accumulated in
    float64 here so that the CUDA result (fp32 tree sums) is compared against a more accurate
    value with a stated tolerance."""
    X = np.asarray(X)
    X = X.reshape(-1, X.shape[-1]).astype(np.float64)
    return (X * X).sum(axis=0)


def scaler_row_update(scaler_row, nsamples: int, X):
    """One WrappedGPT.add_batch call (layerwrapper.py:22-35): X is [b, T, K] or [T, K].
    scaler_row *= n/(n+b); n += b; scaler_row += ||X[:,k]||_2^2 / n.   Returns (new, n)."""
    X = np.asarray(X)
    b = 1 if X.ndim == 2 else X.shape[0]
    sr = np.asarray(scaler_row, dtype=F32) * F32(nsamples / (nsamples + b))
    nsamples += b
    ss = colsumsq(X)
    return (sr + (ss / nsamples).astype(F32)).astype(F32), nsamples


def dead_columns(X) -> np.ndarray:
    """mxqgpt.py:401: dead = diag(H) == 0, H = sum_samples (2/n) X^T X  (mxqgpt.py:377-383)."""
    X = np.asarray(X)
    X = X.reshape(-1, X.shape[-1])
    return ~(X != 0).any(axis=0)


def wanda_metric(W, scaler_row):
    """prune.py:177: |W| * sqrt(scaler_row.reshape(1,-1)) in fp32."""
    W = np.asarray(W, dtype=F32)
    return np.abs(W) * np.sqrt(np.asarray(scaler_row, dtype=F32))[None, :]


# --------------------------------------------------------------------------------------
# (a-7) Quantizer.find_params / quantize_dequantize   quantizer.py:5-20,61-121,149-155
# --------------------------------------------------------------------------------------

def _find_params(xmin, xmax, maxq: int):
    """quantizer.py:81-99 (perchannel, asym, round_zero=False)."""
    xmin = xmin.astype(F32).copy()
    xmax = xmax.astype(F32).copy()
    eq = xmin == xmax                               # :90-92
    xmin[eq] = -1
    xmax[eq] = +1
    scale = ((xmax - xmin) / F32(maxq)).astype(F32)  # :94
    with np.errstate(divide="ignore", invalid="ignore"):
        zero = (-xmin / scale).astype(F32)          # :99
    return scale, zero


def _qq_scale(scale, qq_bits: int = 4, qq_group: int = 16):
    """quantizer.py:114-121: second-level min/max quantization of `scale` over `qq_group`
    consecutive entries along axis 0 (= consecutive output rows).  Returns (dequantized scale,
    integer codes, scale2, zero2)."""
    n = scale.shape[0]
    if n % qq_group:
        raise ValueError("rows must be a multiple of qq_groupsize=16 (quantizer.py:115)")
    maxq = 2 ** qq_bits - 1
    sg = scale.reshape((n // qq_group, qq_group) + scale.shape[1:])
    s2, z2 = _find_params(sg.min(axis=1), sg.max(axis=1), maxq)
    s2b, z2b = s2[:, None], z2[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.clip(np.rint((sg / np.maximum(s2b, F32(1e-9))).astype(F32) + z2b), 0, maxq).astype(F32)
    deq = (s2b * (q - z2b)).astype(F32)             # dequantize(): scale * (x - zero), :19-20
    return deq.reshape(scale.shape), q.reshape(scale.shape).astype(np.uint8), s2, z2


def _quant_dequant(x, scale, zero, maxq: int):
    """quantizer.py:5-7."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.clip(np.rint((x / np.maximum(scale, F32(1e-9))).astype(F32) + zero), 0, maxq).astype(F32)
    return (scale * (q - zero)).astype(F32), q


# --------------------------------------------------------------------------------------
# (a-6) MXQGPT.fasterquant(blocksize=16)   mxq_quant/lib/mxqgpt.py:387-448
# --------------------------------------------------------------------------------------

def fasterquant(W16, dead=None, group: int = 16, low_bits: int = 2,
                group_bits: np.ndarray | None = None, return_aux: bool = False):
    """W16: [N, K] fp16 weights.  Returns the fake-quantized fp16 weights (and aux).

    :401-403  W[:, dead] = 0
    :413-428  every low-bit group of every row: Quantizer(bits=low_bits, perchannel, asym,
              qq_scale_bits=4).find_params + quantize_dequantize
    :431-436  4-bit columns of a row pooled, one Quantizer(bits=4, ..., qq_scale_bits=4)
    :448      cast back to the weight dtype (fp16)
    """
    W = np.asarray(W16, dtype=np.float16).astype(F32)
    N, K = W.shape
    ng = K // group
    if group_bits is None:
        group_bits = reference_group_bits(K, group, low_bits)
    bits, pool = _split_bits(group_bits)
    if (~pool).any() and len(set(bits[~pool].tolist())) != 1:
        raise ValueError("un-pooled groups must share one bit-width")
    if dead is not None:
        W = W.copy()
        W[:, np.asarray(dead, dtype=bool)] = 0
    xg = W.reshape(N, ng, group)
    out = np.empty_like(xg)
    codes = np.empty(xg.shape, dtype=np.uint8)
    low = ~pool
    aux = {}
    if low.any():
        xl = xg[:, low, :]
        maxq = 2 ** int(bits[low][0]) - 1
        scale, zero = _find_params(xl.min(axis=2), xl.max(axis=2), maxq)      # [N, nlow]
        scale_q, scode, s2, z2 = _qq_scale(scale)
        o, q = _quant_dequant(xl, scale_q[:, :, None], zero[:, :, None], maxq)
        out[:, low, :] = o
        codes[:, low, :] = q.astype(np.uint8)
        aux.update(scale_low=scale_q, zero_low=zero, scale_code_low=scode, scale2_low=s2, zero2_low=z2)
    if pool.any():
        xp = xg[:, pool, :]
        flat = xp.reshape(N, -1)
        maxq4 = 2 ** int(bits[pool][0]) - 1
        scale, zero = _find_params(flat.min(axis=1), flat.max(axis=1), maxq4)  # [N]
        scale_q, scode, s2, z2 = _qq_scale(scale)
        o, q = _quant_dequant(xp, scale_q[:, None, None], zero[:, None, None], maxq4)
        out[:, pool, :] = o
        codes[:, pool, :] = q.astype(np.uint8)
        aux.update(scale_4b=scale_q, zero_4b=zero, scale_code_4b=scode, scale2_4b=s2, zero2_4b=z2)
    with np.errstate(over="ignore"):
        W_out = out.reshape(N, K).astype(np.float16)
    if return_aux:
        aux["codes"] = codes.reshape(N, K)
        return W_out, aux
    return W_out


def fasterquant_blocksize(W16, dead=None, blocksize: int = 128):
    """MXQGPT.fasterquant with a `blocksize` other than 16 (the signature's default is 128,
    mxqgpt.py:388): inside every 64-column block the 2-bit columns [0, 48) are cut into groups
    `range(0, 48, blocksize)` (:413-415: 128 -> one 48-wide group, 32 -> 32 + 16), the last 16
    columns of all blocks are pooled per row as before (:431-436)."""
    W = np.asarray(W16, dtype=np.float16).astype(F32).copy()
    N, K = W.shape
    if K % 64:
        raise ValueError("columns must be a multiple of 64")
    if dead is not None:
        W[:, np.asarray(dead, dtype=bool)] = 0
    out = W.copy()
    xb = W.reshape(N, K // 64, 64)
    ob = out.reshape(N, K // 64, 64)
    for j0 in range(0, 48, blocksize):
        j1 = min(j0 + blocksize, 48)
        seg = xb[:, :, j0:j1]
        scale, zero = _find_params(seg.min(axis=2), seg.max(axis=2), 3)
        scale_q = _qq_scale(scale)[0]
        ob[:, :, j0:j1] = _quant_dequant(seg, scale_q[:, :, None], zero[:, :, None], 3)[0]
    pool = xb[:, :, 48:]
    flat = pool.reshape(N, -1)
    scale, zero = _find_params(flat.min(axis=1), flat.max(axis=1), 15)
    scale_q = _qq_scale(scale)[0]
    ob[:, :, 48:] = _quant_dequant(pool, scale_q[:, None, None], zero[:, None, None], 15)[0]
    with np.errstate(over="ignore"):
        return out.astype(np.float16)


# --------------------------------------------------------------------------------------
# (a-9) packed mixed 2/4-bit layout   cuda_kernel/csrc/quantization/gemv_mxq_cuda.cu:39-208
# --------------------------------------------------------------------------------------

def packed_shapes(OC: int, IC: int) -> dict:
    """Tensor shapes of the packed layout.  At IC == 4096 they are the reference's
    (gemv_mxq_cuda.cu:54-59,69-70,135; test_correct_gemv.py:23-36); for other IC the
    per-block metadata words are tiled in chunks of 64 blocks (SURVEY.md 8a 'layout
    generalisation')."""
    if IC % 64 or OC % 8:
        raise ValueError("IC must be a multiple of 64 and OC a multiple of 8")
    nblk = IC // 64
    nchunk = (nblk + 63) // 64
    return dict(weight=(OC, nblk * 4), weight_last=(OC, nblk), zeros_and_scales=(OC, 32 * nchunk),
                zeros_2nd=(OC // 4, 32 * nchunk), scales_2nd=(OC // 4, 3 * nblk),
                scales_4b=(OC,), zeros_4b=(OC // 8,))


def _meta_pos(nblk: int):
    b = np.arange(nblk)
    word = (b // 64) * 32 + (b % 64) % 32
    half = (b % 64) // 32
    return word, half


def decode_mxq(p: dict, IC: int | None = None) -> np.ndarray:
    """Dequantize a packed tensor set to fp32 [OC, IC] exactly as the reference kernel does:
      2-bit: scaling = float(s2[oc/4, 3*blk+k]) * (c - z2);  w = scaling * (q - z1)
             (gemv_mxq_cuda.cu:131-136,152-153)
      4-bit: w = float(s4[oc]) * (q - z4)                     (:61-62,178-179,191-192)
    with the bit positions of :101-103,109-110,144-159,164,176-199."""
    weight = np.asarray(p["weight"]).view(np.uint32)
    wlast = np.asarray(p["weight_last"]).view(np.uint32)
    zs = np.asarray(p["zeros_and_scales"]).view(np.uint32)
    z2w = np.asarray(p["zeros_2nd"]).view(np.uint32)
    s2 = np.asarray(p["scales_2nd"], dtype=np.float16).astype(F32)
    s4 = np.asarray(p["scales_4b"], dtype=np.float16).astype(F32)
    z4w = np.asarray(p["zeros_4b"]).view(np.uint32)
    OC = weight.shape[0]
    nblk = weight.shape[1] // 4
    word, half = _meta_pos(nblk)
    out = np.empty((OC, nblk, 64), dtype=F32)
    oc = np.arange(OC)
    j16 = np.arange(16, dtype=np.uint32)
    j8 = np.arange(8, dtype=np.uint32)
    zbyte = (zs[:, word] >> (16 * half)[None, :].astype(np.uint32)) & np.uint32(0xFF)        # [OC, nblk]
    cbyte = (zs[:, word] >> (16 * half + 8)[None, :].astype(np.uint32)) & np.uint32(0xFF)
    z2byte = (z2w[:, word] >> (8 * half)[None, :].astype(np.uint32)) & np.uint32(0xFF)       # [OC/4, nblk]
    z2byte = z2byte[oc // 4]
    for k in range(3):
        z1 = ((zbyte >> np.uint32(2 * k)) & np.uint32(3)).astype(F32)
        c = ((cbyte >> np.uint32(2 * k)) & np.uint32(3)).astype(F32)
        z2 = ((z2byte >> np.uint32(2 * k)) & np.uint32(3)).astype(F32)
        scaling = (s2[oc // 4][:, k::3] * (c - z2)).astype(F32)                               # [OC, nblk]
        q = ((weight[:, k::4][:, :, None] >> (2 * j16)[None, None, :]) & np.uint32(3)).astype(F32)
        out[:, :, 16 * k:16 * k + 16] = (scaling[:, :, None] * (q - z1[:, :, None])).astype(F32)
    z4 = ((z4w[oc // 8] >> (4 * (oc % 8)).astype(np.uint32)) & np.uint32(0xF)).astype(F32)
    q = ((weight[:, 3::4][:, :, None] >> (4 * j8)[None, None, :]) & np.uint32(0xF)).astype(F32)
    out[:, :, 48:56] = (s4[:, None, None] * (q - z4[:, None, None])).astype(F32)
    q = ((wlast[:, :, None] >> (4 * j8)[None, None, :]) & np.uint32(0xF)).astype(F32)
    out[:, :, 56:64] = (s4[:, None, None] * (q - z4[:, None, None])).astype(F32)
    out = out.reshape(OC, nblk * 64)
    if IC is not None:
        assert out.shape[1] == IC
    return out


def gemv_mxq(x16, p: dict) -> np.ndarray:
    """y[b, oc] = sum_k decode(W)[oc, k] * x[b, k], fp32 accumulate (float64 here, compared
    with tolerance), fp16 output (gemv_mxq_cuda.cu:202-206)."""
    Wd = decode_mxq(p).astype(np.float64)
    x = np.asarray(x16, dtype=np.float16).astype(np.float64)
    with np.errstate(over="ignore"):
        return (x @ Wd.T).astype(np.float16)


def gemm_mxq_f32(x16, p: dict) -> np.ndarray:
    """fp32-accurate reference for the prefill GEMM: x[M, IC] @ decode(W)^T in float64."""
    Wd = decode_mxq(p).astype(np.float64)
    x = np.asarray(x16, dtype=np.float16).astype(np.float64)
    return x @ Wd.T


def pack_mxq(W16, dead=None) -> dict:
    """Quantize fp16 weights [OC, IC] into the packed layout.  ENCODE POLICY IS OURS
    (parity unpinned, see header).  All arithmetic fp32, op by op:

    2-bit group (row oc, block b, slot k<3), lo = min(min(w),0), hi = max(max(w),0):
        s  = (hi - lo) / 3                       (s == 0 -> 1)
        second level over the 4 rows oc//4*4..+3 of the same (b,k):
        s2 = fp16(max_rows(s) / 3), z2 = 0, c = clamp(rint(s / float(s2)), 1, 3)
        S  = float(s2) * c ; z1 = clamp(rint(-lo / S), 0, 3)
        q  = clamp(rint(w / S) + z1, 0, 3)
    4-bit pool of a row: lo/hi over all 4-bit columns of the row (zero included)
        s4 = fp16((hi - lo) / 15) (0 -> 1); z4 = clamp(rint(-lo / float(s4)), 0, 15)
        q  = clamp(rint(w / float(s4)) + z4, 0, 15)
    """
    W = np.asarray(W16, dtype=np.float16).astype(F32)
    OC, IC = W.shape
    shp = packed_shapes(OC, IC)
    if dead is not None:
        W = W.copy()
        W[:, np.asarray(dead, dtype=bool)] = 0
    nblk = IC // 64
    xb = W.reshape(OC, nblk, 4, 16)
    x2 = xb[:, :, 0:3, :]
    lo = np.minimum(x2.min(axis=3), F32(0))
    hi = np.maximum(x2.max(axis=3), F32(0))
    s = ((hi - lo) / F32(3)).astype(F32)
    s[s == 0] = 1
    smax = s.reshape(OC // 4, 4, nblk, 3).max(axis=1)
    with np.errstate(over="ignore"):
        s2 = (smax / F32(3)).astype(F32).astype(np.float16)                     # [OC/4, nblk, 3]
    s2f = np.repeat(s2.astype(F32), 4, axis=0)                                  # [OC, nblk, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.clip(np.rint((s / s2f).astype(F32)), 1, 3).astype(F32)
        S = (s2f * c).astype(F32)
        z1 = np.clip(np.rint((-lo / S).astype(F32)), 0, 3).astype(F32)
        q2 = np.clip(np.rint((x2 / S[..., None]).astype(F32)) + z1[..., None], 0, 3).astype(np.uint32)
    x4 = xb[:, :, 3, :]                                                          # [OC, nblk, 16]
    lo4 = np.minimum(x4.reshape(OC, -1).min(axis=1), F32(0))
    hi4 = np.maximum(x4.reshape(OC, -1).max(axis=1), F32(0))
    with np.errstate(over="ignore"):
        s4 = ((hi4 - lo4) / F32(15)).astype(F32).astype(np.float16)
    s4[s4 == 0] = 1
    s4f = s4.astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        z4 = np.clip(np.rint((-lo4 / s4f).astype(F32)), 0, 15).astype(F32)
        q4 = np.clip(np.rint((x4 / s4f[:, None, None]).astype(F32)) + z4[:, None, None], 0, 15).astype(np.uint32)

    j16 = (2 * np.arange(16, dtype=np.uint32))
    j8 = (4 * np.arange(8, dtype=np.uint32))
    weight = np.zeros(shp["weight"], dtype=np.uint32)
    for k in range(3):
        weight[:, k::4] = np.bitwise_or.reduce(q2[:, :, k, :] << j16[None, None, :], axis=2)
    weight[:, 3::4] = np.bitwise_or.reduce(q4[:, :, 0:8] << j8[None, None, :], axis=2)
    weight_last = np.bitwise_or.reduce(q4[:, :, 8:16] << j8[None, None, :], axis=2).astype(np.uint32)

    word, half = _meta_pos(nblk)
    zbyte = np.zeros((OC, nblk), dtype=np.uint32)
    cbyte = np.zeros((OC, nblk), dtype=np.uint32)
    for k in range(3):
        zbyte |= z1[:, :, k].astype(np.uint32) << np.uint32(2 * k)
        cbyte |= c[:, :, k].astype(np.uint32) << np.uint32(2 * k)
    hw = zbyte | (cbyte << np.uint32(8))                                        # 16-bit half-word
    zs = np.zeros(shp["zeros_and_scales"], dtype=np.uint32)
    for p in range(2):
        sel = half == p
        np.bitwise_or.at(zs, (slice(None), word[sel]), hw[:, sel] << np.uint32(16 * p))
    zeros_2nd = np.zeros(shp["zeros_2nd"], dtype=np.uint32)                      # z2 == 0 policy
    scales_2nd = s2.reshape(OC // 4, nblk * 3)
    zeros_4b = np.zeros(shp["zeros_4b"], dtype=np.uint32)
    oc = np.arange(OC)
    np.bitwise_or.at(zeros_4b, oc // 8, z4.astype(np.uint32) << (4 * (oc % 8)).astype(np.uint32))
    return dict(weight=weight.view(np.int32), weight_last=weight_last.view(np.int32),
                zeros_and_scales=zs.view(np.int32), zeros_2nd=zeros_2nd.view(np.int32),
                scales_2nd=np.ascontiguousarray(scales_2nd), scales_4b=s4,
                zeros_4b=zeros_4b.view(np.int32))


def kat_constant_fill(OC: int = 4096, IC: int = 4096) -> dict:
    """The reference's known-answer vector (cuda_kernel/test_correct_gemv.py:19-40): every
    dequantized weight equals 1, so y == IC for x == 1."""
    shp = packed_shapes(OC, IC)
    u = lambda v: np.uint32(v).astype(np.uint32).view(np.int32)
    return dict(weight=np.full(shp["weight"], u(0xAAAAAAAA), dtype=np.int32),
                weight_last=np.full(shp["weight_last"], u(0xAAAAAAAA), dtype=np.int32),
                zeros_and_scales=np.full(shp["zeros_and_scales"], u(0xAA55AA55), dtype=np.int32),
                zeros_2nd=np.full(shp["zeros_2nd"], u(0x55555555), dtype=np.int32),
                scales_2nd=np.ones(shp["scales_2nd"], dtype=np.float16),
                scales_4b=np.ones(shp["scales_4b"], dtype=np.float16),
                zeros_4b=np.full(shp["zeros_4b"], u(0x99999999), dtype=np.int32))


def random_packed(OC: int, IC: int, seed: int = 0) -> dict:
    """Raw random-bit packed tensors with s2, s4 ~ U(0.001, 0.01)."""
    rng = np.random.default_rng(seed)
    shp = packed_shapes(OC, IC)
    ri = lambda s: rng.integers(0, 2 ** 32, size=s, dtype=np.uint64).astype(np.uint32).view(np.int32)
    return dict(weight=ri(shp["weight"]), weight_last=ri(shp["weight_last"]),
                zeros_and_scales=ri(shp["zeros_and_scales"]), zeros_2nd=ri(shp["zeros_2nd"]),
                scales_2nd=rng.uniform(0.001, 0.01, shp["scales_2nd"]).astype(np.float16),
                scales_4b=rng.uniform(0.001, 0.01, shp["scales_4b"]).astype(np.float16),
                zeros_4b=ri(shp["zeros_4b"]))
