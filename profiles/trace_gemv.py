"""Per-phase timeline of the decode GEMV inside a CUDA-graph chain (MXQ_GEMV_DBG=8 stamps)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import _lib as L, ops  # noqa: E402
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


os.environ["MXQ_GEMV_DBG"] = "8"
for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
    nset = 24
    ps = [rand_packed(oc, ic) for _ in range(nset)]
    x = torch.randn(1, ic, device=dev).half()
    y = torch.empty(1, oc, device=dev, dtype=torch.float16)
    for pdl in (True, False):
        for p in ps[:2]:
            ops.gemv(x, p, out=y, validate=False, pdl=pdl)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for p in ps:
                ops.gemv(x, p, out=y, validate=False, pdl=pdl)
        g.replay()
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 640)()
        L.lib().mxq_debug_gemv_trace(buf)
        t = np.array(buf, dtype=np.int64).reshape(160, 4)
        t = t[t[:, 0] > 0]
        ref = t[:, 0].min()
        t = t - ref
        print(f"{oc}x{ic} pdl={int(pdl)} last kernel of the chain, {len(t)} CTAs, ns rel. to first CTA start:")
        for i, name in enumerate(("start", "waited", "staged", "done")):
            print(f"   {name:7s} min {t[:, i].min():6d}  mean {t[:, i].mean():8.0f}  max {t[:, i].max():6d}")

print("standalone launches (same weights twice, L2 warm, no graph):")
for oc, ic in ((4096, 4096), (4096, 11008)):
    p = rand_packed(oc, ic)
    x = torch.randn(1, ic, device=dev).half()
    y = torch.empty(1, oc, device=dev, dtype=torch.float16)
    for pdl in (False, True):
        for _ in range(3):
            ops.gemv(x, p, out=y, validate=False, pdl=pdl)
            torch.cuda.synchronize()
        buf = (C.c_ulonglong * 640)()
        L.lib().mxq_debug_gemv_trace(buf)
        t = np.array(buf, dtype=np.int64).reshape(160, 4)[:147]
        t = t - t[:, 0].min()
        print(f"{oc}x{ic} pdl={int(pdl)}: " + "  ".join(f"{n} {t[:, i].mean():.0f}" for i, n in enumerate(("start", "waited", "staged", "done"))) + f"  done.max {t[:, 3].max()}")
