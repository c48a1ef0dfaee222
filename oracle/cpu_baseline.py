"""CPU baseline legs for bench.py (cpu_baseline / --impl reference) -- TEST INFRASTRUCTURE.

The reference's Python cannot travel to the GPU box (/root/reference is absent there), so the CPU
arm times the oracle port (kind = "port") of the same algorithms, threaded over independent row
slabs with all host cores (numpy releases the GIL inside its kernels).  Note that the real
reference additionally spends 68.7 GFLOP per 2048-token sample per 4096-wide linear on a K x K
Hessian it never uses (mxqgpt.py:383); the port only computes the diagonal, so this baseline is
faster than the reference itself would be.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import mxq_oracle as O


def _slabs(n: int, step: int):
    return [(i, min(i + step, n)) for i in range(0, n, step)]


def _pmap(fn, items, threads):
    if threads <= 1:
        return [fn(it) for it in items]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, items))


def colsumsq_threaded(X: np.ndarray, threads: int) -> np.ndarray:
    parts = _pmap(lambda s: O.colsumsq(X[s[0]:s[1]]), _slabs(X.shape[0], 1024), threads)
    return np.sum(parts, axis=0)


def fasterquant_threaded(W: np.ndarray, dead, threads: int) -> np.ndarray:
    out = np.empty_like(W)

    def run(s):
        out[s[0]:s[1]] = O.fasterquant(W[s[0]:s[1]], dead)
    _pmap(run, _slabs(W.shape[0], 256), threads)
    return out


def pack_threaded(W: np.ndarray, dead, threads: int):
    return _pmap(lambda s: O.pack_mxq(W[s[0]:s[1]], dead), _slabs(W.shape[0], 256), threads)


def fakequant_threaded(x: np.ndarray, dtype: str, threads: int) -> np.ndarray:
    out = np.empty_like(x)

    def run(s):
        out[s[0]:s[1]] = O.fakequant_fwd(x[s[0]:s[1]], dtype)
    _pmap(run, _slabs(x.shape[0], 128), threads)
    return out


def ptq_layer_sample(hidden: int, inter: int, tokens: int, threads: int | None = None,
                     row_fraction: float = 1.0, seed: int = 0):
    """One decoder layer of the mxq pass on the CPU: 4 statistics over `tokens` calibration tokens
    + fasterquant + pack of the 7 linears (optionally only the first `row_fraction` of each
    linear's rows, 16-row aligned).  Returns (seconds, algorithmic_bytes, description)."""
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    shapes = [(hidden, hidden, "a"), (hidden, hidden, "a"), (hidden, hidden, "a"), (hidden, hidden, "o"),
              (inter, hidden, "m"), (inter, hidden, "m"), (hidden, inter, "d")]
    calib = {k: rng.standard_normal((tokens, d), dtype=np.float32).astype(np.float16)
             for k, d in (("a", hidden), ("o", hidden), ("m", hidden), ("d", inter))}
    for X in calib.values():
        X[:, 7] = 0
    weights = []
    for oc, ic, key in shapes:
        rows = max(16, int(oc * row_fraction) // 16 * 16)
        weights.append(((rng.standard_normal((rows, ic), dtype=np.float32) * 0.02).astype(np.float16), key))
    t0 = time.perf_counter()
    stats = {k: colsumsq_threaded(X, threads) for k, X in calib.items()}
    nbytes = sum(X.size * 2 for X in calib.values())
    for W, key in weights:
        dead = stats[key] == 0
        fasterquant_threaded(W, dead, threads)
        pack_threaded(W, dead, threads)
        nbytes += W.size * 4 + int(W.size * 0.3756)
    dt = time.perf_counter() - t0
    desc = (f"1 decoder layer ({hidden}/{inter}), {tokens} calibration tokens, "
            f"{row_fraction:.3g} of each linear's rows, numpy oracle port on {threads} threads")
    return dt, nbytes, desc


def fakequant_sample(rows: int, cols: int, dtype: str, threads: int | None = None, seed: int = 0):
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, cols), dtype=np.float32) * 0.02)
    x = O.rounder(dtype)(x)
    g = rng.standard_normal((rows, cols), dtype=np.float32)
    t0 = time.perf_counter()
    fakequant_threaded(x, dtype, threads)
    O.ste_bwd(g, x)
    dt = time.perf_counter() - t0
    esz = 4 if dtype == "fp32" else 2
    return dt, rows * cols * esz * 5, f"{rows}x{cols} {dtype} fwd+bwd, numpy oracle port on {threads} threads"


def dequant_matmul_sample(rows: int, cols: int, M: int, threads: int | None = None, what: str = "bytes", seed: int = 0):
    """CPU leg of the packed dequant-GEMV / GEMM (the reference ships no CPU path for the packed
    format): restated decode of gemv_mxq_cuda.cu:131-136,152-153,178-179 on `rows` output rows
    (threaded over row slabs) followed by an fp32 matmul with an [M, cols] activation block.
    Returns (seconds, work, description); work = packed bytes (+x, y) or 2*M*rows*cols FLOP."""
    threads = threads or os.cpu_count() or 1
    p = O.random_packed(rows, cols, seed=seed)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((M, cols), dtype=np.float32).astype(np.float16)
    slabs = _slabs(rows, 64)

    def sub(s):
        r0, r1 = s
        q = dict(weight=p["weight"][r0:r1], weight_last=p["weight_last"][r0:r1], zeros_and_scales=p["zeros_and_scales"][r0:r1],
                 zeros_2nd=p["zeros_2nd"][r0 // 4:r1 // 4], scales_2nd=p["scales_2nd"][r0 // 4:r1 // 4],
                 scales_4b=p["scales_4b"][r0:r1], zeros_4b=p["zeros_4b"][r0 // 8:r1 // 8])
        return x.astype(np.float32) @ O.decode_mxq(q).astype(np.float32).T
    t0 = time.perf_counter()
    _pmap(sub, slabs, threads)
    dt = time.perf_counter() - t0
    nbytes = sum(v.nbytes for v in p.values()) + 2 * M * (rows + cols)
    work = nbytes if what == "bytes" else 2.0 * M * rows * cols
    return dt, work, (f"decode + fp32 matmul of a packed {rows}x{cols} linear, M = {M}, numpy oracle port on {threads} threads")
