"""Phase probe of the decode GEMV: time a CUDA-graph chain with phases disabled (MXQ_GEMV_DBG)."""
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


def empty_chain(n, pdl):
    a = torch.zeros(8, device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            a.add_(1.0)
    return g


for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
    nset = max(4, int(400e6 / packed_nbytes(oc, ic)))
    ps = [rand_packed(oc, ic) for _ in range(nset)]
    x = torch.randn(1, ic, device=dev).half()
    y = torch.empty(1, oc, device=dev, dtype=torch.float16)
    os.environ["MXQ_GEMV_VERBOSE"] = "1"
    ops.gemv(x, ps[0], out=y, validate=False)
    os.environ.pop("MXQ_GEMV_VERBOSE")
    for dbg in ("0", "1", "2", "3"):
        for pdl in (True, False):
            os.environ["MXQ_GEMV_DBG"] = dbg
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
            nb = packed_nbytes(oc, ic)
            print(f"{oc}x{ic} dbg={dbg} pdl={int(pdl)}: {us:.2f} us/gemv = {nb / us / 1e3:.0f} GB/s")
os.environ["MXQ_GEMV_DBG"] = "0"
g = empty_chain(200, False)
g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"torch add_ chain: {a.elapsed_time(b) / 5 / 200 * 1e3:.2f} us/kernel")
