"""Sweep the fake-quant launch knobs (stages, CTAs/SM, team bytes) on the BASELINE q_proj shape."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
lib = ops.L.lib()


def timeit(fn, iters=40):
    for _ in range(5):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for dt, shape in ((torch.float32, (4096, 4096)), (torch.bfloat16, (4096, 4096)), (torch.bfloat16, (4096, 11008)),
                  (torch.float32, (4096, 11008))):
    nset = 6
    xs = [(torch.randn(*shape, device=dev) * 0.02).to(dt) for _ in range(nset)]
    outs = [torch.empty_like(xs[0]) for _ in range(nset)]
    nb = shape[0] * shape[1] * xs[0].element_size() * 2

    def fwd(i):
        k = i % nset
        lib.mxq_fakequant_fwd(xs[k].data_ptr(), outs[k].data_ptr(), None, shape[0], shape[1],
                              ops.L.dtype_enum(xs[k]), 16, 2, None, ops.L.stream())

    def copy(i):
        k = i % nset
        outs[k].copy_(xs[k])
    t = timeit(copy)
    print(f"{dt} {shape}: torch copy_ {t:.1f} us = {nb / t / 1e3:.0f} GB/s")
    for stages, bps, tb in itertools.product((2, 3, 4), (2, 3, 4, 5), (16384, 32768)):
        os.environ["MXQ_FQ_STAGES"] = str(stages)
        os.environ["MXQ_FQ_BPS"] = str(bps)
        os.environ["MXQ_FQ_TEAM_BYTES"] = str(tb)
        t = timeit(fwd)
        print(f"  stages={stages} bps={bps} team_bytes={tb}: {t:.1f} us = {nb / t / 1e3:.0f} GB/s")
    for k in ("MXQ_FQ_STAGES", "MXQ_FQ_BPS", "MXQ_FQ_TEAM_BYTES"):
        os.environ.pop(k, None)
