"""ncu target: fused fasterquant + pack on 4096^2 and 4096x11008."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200.prune import LinearQuantJob  # noqa: E402

dev = torch.device("cuda:0")
for oc, ic in ((4096, 4096), (4096, 11008)):
    W = (torch.randn(oc, ic, device=dev) * 0.02).half()
    stat = torch.rand(ic, device=dev) + 0.1
    stat[7] = 0
    job = LinearQuantJob(oc, ic, dev)
    for _ in range(2):
        job.run(W, stat)
    torch.cuda.synchronize()
