"""Decode GEMV timing per shape (CUDA graph of one GEMV per distinct weight set, > L2), with and
without programmatic dependent launch.  MXQ_GEMV_WARPS / MXQ_GEMV_WPR / MXQ_GEMV_STAGES override the plan."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402
from mxq_b200.prune import packed_nbytes  # noqa: E402

dev = torch.device("cuda:0")


def rand_packed(oc, ic):
    p = {}
    for k, (s, d) in ops.packed_shapes(oc, ic).items():
        if d == torch.float16:
            p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
        else:
            p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
    return p


for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
    nset = max(4, int(400e6 / packed_nbytes(oc, ic)))
    ps = [rand_packed(oc, ic) for _ in range(nset)]
    for B in (1, 4):
        x = torch.randn(B, ic, device=dev).half()
        y = torch.empty(B, oc, device=dev, dtype=torch.float16)
        for ks in ("auto", "auto-nopdl"):
            pdl = ks != "auto-nopdl"
            if ks.startswith("auto"):
                os.environ.pop("MXQ_GEMV_WPR", None)
            else:
                os.environ["MXQ_GEMV_WPR"] = ks
            for p in ps[:2]:
                ops.gemv(x, p, out=y, validate=False, pdl=pdl)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for p in ps:
                    ops.gemv(x, p, out=y, validate=False, pdl=pdl)
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 5 / nset * 1e3
            nb = packed_nbytes(oc, ic) + 2 * B * (oc + ic)
            print(f"{oc}x{ic} B={B} WPR={ks}: {us:.2f} us/gemv = {nb / us / 1e3:.0f} GB/s")
