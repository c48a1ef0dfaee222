"""Time the activation / KV-cache fake quantizers (SymQuantizer, AsymQuantizer) on QAT-sized tensors
with rotating buffers larger than L2, next to torch copy_ of the same tensors."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


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


CASES = [
    ("sym", 8, torch.bfloat16, (2, 2048, 4096), False),     # QuantizeLinear input, run_train.sh 2 8 x
    ("sym", 8, torch.bfloat16, (2, 2048, 11008), False),    # down_proj input
    ("asym", 4, torch.bfloat16, (2, 2048, 4096), False),    # KV cache, kv_bits = 4
    ("sym", 8, torch.float32, (2, 2048, 4096), False),
    ("sym", 8, torch.bfloat16, (4096, 4096), False),        # 2-D: 128-column groups
    ("asym", 4, torch.bfloat16, (4096, 4096), False),       # 2-D: 8-column groups
    ("asym", 4, torch.bfloat16, (2, 32, 2048, 128), False),  # 4-D per head
    ("sym", 8, torch.bfloat16, (2, 2048, 4096), True),      # layerwise
]
for mode, bits, dt, shape, lw in CASES:
    nset = 6
    xs = [torch.randn(*shape, device=dev).to(dt) for _ in range(nset)]
    outs = [torch.empty_like(xs[0]) for _ in range(nset)]
    nseg, seglen, period, valid = ops.segquant_plan(shape, mode, lw)
    need = ops.L.lib().mxq_segquant_workspace_bytes(nseg, seglen, ops.L.dtype_enum(xs[0]))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
    nb = xs[0].numel() * xs[0].element_size() * 2
    lib = ops.L.lib()

    def fwd(i):
        k = i % nset
        rc = lib.mxq_segquant_fwd(xs[k].data_ptr(), outs[k].data_ptr(), nseg, seglen,
                                  ops.L.dtype_enum(xs[k]), 0 if mode == "sym" else 1, bits, period, valid,
                                  ws.data_ptr(), ws.numel(), ops.L.stream())
        assert rc == 0, rc

    def copy(i):
        k = i % nset
        outs[k].copy_(xs[k])
    tc = timeit(copy)
    t = timeit(fwd)
    print(f"{mode}{bits} {str(dt)[6:]} {shape} lw={lw}: {t:.1f} us = {nb / t / 1e3:.0f} GB/s "
          f"(copy_ {tc:.1f} us = {nb / tc / 1e3:.0f} GB/s; segments {nseg} x {seglen})")
