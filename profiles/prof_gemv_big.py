"""Steady-state decode GEMV: one very tall matrix (65536 x 4096, 100 MB packed) so that the main
loop dominates launch, staging and tail; for ncu source-level stall attribution."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402
from mxq_b200.prune import packed_nbytes  # noqa: E402

dev = torch.device("cuda:0")
oc, ic = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = {}
for k, (s, d) in ops.packed_shapes(oc, ic).items():
    if d == torch.float16:
        p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
    else:
        p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
x = torch.randn(1, ic, device=dev).half()
y = torch.empty(1, oc, device=dev, dtype=torch.float16)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for _ in range(3):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.gemv(x, p, out=y, validate=False)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3
    print(f"{oc}x{ic}: {us:.1f} us = {packed_nbytes(oc, ic) / us / 1e3:.0f} GB/s")
