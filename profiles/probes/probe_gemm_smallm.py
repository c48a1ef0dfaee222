"""Prefill / batched-decode GEMM at M <= 256: single-CTA kernel (OC / 256 tiles, MXQ_GEMM_SINGLE=1)
vs the CTA-pair kernel with K-split tiles, vs cuBLAS fp16; CUDA graph of 20 launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=20):
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
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


for OC, IC in ((4096, 4096), (11008, 4096), (4096, 11008)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    p = ops.pack(W)
    for M in (16, 64, 128, 256):
        x = torch.randn(M, IC, device=dev).half()
        y = torch.empty(M, OC, device=dev, dtype=torch.float16)
        ref = x.float() @ ops.unpack(p).T
        os.environ["MXQ_GEMM_SINGLE"] = "1"
        ws = ops.gemm_workspace(M, IC, OC, dev)
        t1 = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
        e1 = float((y.float() - ref).abs().max() / ref.abs().max())
        os.environ.pop("MXQ_GEMM_SINGLE")
        ws = ops.gemm_workspace(M, IC, OC, dev)
        t2 = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
        e2 = float((y.float() - ref).abs().max() / ref.abs().max())
        tc = timeit(lambda: torch.matmul(x, W.t(), out=y))
        print(f"{OC}x{IC} M={M}: single-CTA {t1:.1f} us (err {e1:.1e}) | pair + K split {t2:.1f} us (err {e2:.1e}, ws {ws.numel() >> 20} MiB) | cuBLAS {tc:.1f} us", flush=True)
    del W, p
