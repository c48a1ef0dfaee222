"""Per-phase timeline of the IMMA decode GEMV inside a CUDA-graph chain (MXQ_GEMV_DBG=8 stamps)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda:0")
NAMES = ("start", "issued", "waited", "staged", "loopdone", "done")


def rand_packed(oc, ic):
    p = {}
    for k, (s, d) in ops.packed_shapes(oc, ic).items():
        if d == torch.float16:
            p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
        else:
            p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
    return p


DBG = sys.argv[1] if len(sys.argv) > 1 else "8"
MID = int(sys.argv[2]) if len(sys.argv) > 2 else 12
os.environ["MXQ_GEMV_DBG"] = DBG
print("MXQ_GEMV_DBG =", DBG, "traced kernel index", MID)
os.environ["MXQ_GEMV_IMPL"] = "mma"
for oc, ic in ((4096, 4096), (4096, 11008)):
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
            for i, p in enumerate(ps):
                # stamps from ONE kernel in the middle of the chain (it has a predecessor and a successor)
                os.environ["MXQ_GEMV_DBG"] = DBG if i == MID else str(int(DBG) & ~8)
                ops.gemv(x, p, out=y, validate=False, pdl=pdl)
        g.replay()
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 960)()
        L.lib().mxq_debug_gemv2_trace(buf)
        t = np.array(buf, dtype=np.int64).reshape(160, 6)
        t = t[t[:, 0] > 0]
        t = t - t[:, 0].min()
        print(f"{oc}x{ic} pdl={int(pdl)} kernel {MID} of the chain, {len(t)} CTAs, ns rel. to first CTA start:")
        for i, name in enumerate(NAMES):
            print(f"   {name:8s} min {t[:, i].min():6d}  mean {t[:, i].mean():8.0f}  max {t[:, i].max():6d}")
        if int(DBG) & 64:
            ib = (C.c_longlong * (8 * 16 * 6))()
            L.lib().mxq_debug_gemv2_itrace(ib)
            it = np.array(ib, dtype=np.int64).reshape(8, 16, 6)
            for w in (0, 3, 7):
                base = it[w, 0, 0]
                print(f"   warp {w} per-iteration cycles (top, waited, loaded, computed, issued, end) rel. to its first iteration:")
                for k in range(5):
                    if it[w, k, 0] > 0 and it[w, k, 0] >= base:
                        print("      ", [int(v - base) for v in it[w, k]])
