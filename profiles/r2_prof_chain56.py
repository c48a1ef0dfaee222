"""bench.py's decode chain for ncu: the 56 packed linears of 8 Llama-2-7B layers, four activation vectors per
layer, ONE persistent launch (three warm launches, then the one to look at)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from profiles.r2_gemv_persistent import rand_packed, dev  # noqa: E402

shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
xk = [0, 0, 0, 1, 2, 2, 3]
jobs = []
for _ in range(8):
    xl = [torch.randn(ic, device=dev).half() for ic in (4096, 4096, 4096, 11008)]
    for i, (oc, ic) in enumerate(shapes):
        jobs.append((xl[xk[i]], rand_packed(oc, ic), torch.empty(oc, device=dev, dtype=torch.float16), -1))
c = ops.GemvChain(jobs, validate=False)
for _ in range(4):
    c.run()
torch.cuda.synchronize()
