"""ncu target: eager decode GEMVs over distinct packed weights (4096^2, 11008x4096, 4096x11008)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402

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
    ps = [rand_packed(oc, ic) for _ in range(6)]
    x = torch.randn(1, ic, device=dev).half()
    y = torch.empty(1, oc, device=dev, dtype=torch.float16)
    for p in ps:
        ops.gemv(x, p, out=y, validate=False, pdl=False)
    torch.cuda.synchronize()
