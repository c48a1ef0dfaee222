"""A/B: the same GEMM shapes timed (CUDA graph of 20 launches, best of 5 replays) with the library
given in MXQ_AB_LIB (default: the in-tree one), one process per library on the same box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxq_b200 import _lib  # noqa: E402

if os.environ.get("MXQ_AB_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MXQ_AB_LIB"])
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


print("lib:", _lib.LIB_PATH, "split cap:", os.environ.get("MXQ_GEMM_SPLIT"))
for M, OC, IC in ((2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (2048, 8192, 8192), (512, 4096, 4096)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    p = ops.pack(W)
    x = torch.randn(M, IC, device=dev).half()
    y = torch.empty(M, OC, device=dev, dtype=torch.float16)
    fl = 2.0 * M * OC * IC
    need = max(int(_lib.lib().mxq_gemm_workspace_bytes(M, IC, OC)), 4096)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    t = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
    t_c = timeit(lambda: torch.matmul(x, W.t(), out=y))
    print(f"{OC}x{IC} M={M}: {t:.1f} us = {fl / t / 1e6:.0f} TF | cuBLAS {t_c:.1f} us = {fl / t_c / 1e6:.0f} TF", flush=True)
    del W, p
