"""Prefill GEMM tile schedule: whole tiles only (MXQ_GEMM_SPLIT=1) vs K-split tail tiles, next to
cuBLAS fp16, M = 2048 (and a short-M case), Llama-2-7B and 70B shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=20):
    """20 launches captured in one CUDA graph (the Python call path costs more than a short GEMM)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


for M, OC, IC in ((2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (2048, 8192, 8192),
                  (2048, 28672, 8192), (2048, 8192, 28672), (512, 4096, 4096), (1024, 11008, 4096)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    p = ops.pack(W)
    x = torch.randn(M, IC, device=dev).half()
    y = torch.empty(M, OC, device=dev, dtype=torch.float16)
    fl = 2.0 * M * OC * IC
    res = []
    for cap in ("1", "2", "3", "8"):
        os.environ["MXQ_GEMM_SPLIT"] = cap
        ws = ops.gemm_workspace(M, IC, OC, dev)
        t = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
        res.append(f"split<={cap}: {t:.1f} us = {fl / t / 1e6:.0f} TF (ws {ws.numel() >> 20} MiB)")
    os.environ.pop("MXQ_GEMM_SPLIT")
    t_c = timeit(lambda: torch.matmul(x, W.t(), out=y))
    print(f"{OC}x{IC} M={M}: " + " | ".join(res) + f" | cuBLAS {t_c:.1f} us = {fl / t_c / 1e6:.0f} TF", flush=True)
    del W, p
