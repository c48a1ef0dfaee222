"""Prefill GEMM: packed dequant-GEMM vs the same tcgen05 pipeline with a dense TMA-loaded B operand
vs cuBLAS fp16 (torch.matmul), M = 2048, Llama-2-7B and 70B shapes.  Separates the cost of the
in-kernel dequantization from the cost of the pipeline structure."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


M = 2048
for OC, IC in ((4096, 4096), (11008, 4096), (4096, 11008), (8192, 8192), (28672, 8192)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    p = ops.pack(W)
    x = torch.randn(M, IC, device=dev).half()
    y = torch.empty(M, OC, device=dev, dtype=torch.float16)
    ws = torch.zeros(1024, dtype=torch.uint8, device=dev)
    fl = 2.0 * M * OC * IC
    t_p = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
    t_d = timeit(lambda: ops.gemm_dense(x, W))
    t_c = timeit(lambda: torch.matmul(x, W.t(), out=y))
    print(f"{OC}x{IC} M={M}: packed {t_p:.1f} us = {fl / t_p / 1e6:.0f} TF | dense-B pipeline {t_d:.1f} us = "
          f"{fl / t_d / 1e6:.0f} TF | cuBLAS {t_c:.1f} us = {fl / t_c / 1e6:.0f} TF", flush=True)
