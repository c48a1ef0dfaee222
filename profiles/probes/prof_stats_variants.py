"""One launch of each statistics-kernel variant on a 128 x 2048-token, 4096-channel fp16 input (2.1 GB):
register kernel at full occupancy (ctas_per_sm = 8), 16 loads in flight at two resident CTAs (3),
shared-memory TMA ring (0).  For `ncu --set full -k regex:colsumsq`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
L = ops.L
tokens, cols = 128 * 2048, 4096
X = torch.randn(tokens, cols, device=dev, dtype=torch.float16)
out = torch.empty(cols, dtype=torch.float32, device=dev)
ws = torch.empty(L.lib().mxq_colsumsq_workspace_bytes(tokens, cols), dtype=torch.uint8, device=dev)
ref = None
for ctas in (8, 3, 0):
    for _ in range(2):
        rc = L.lib().mxq_colsumsq_ex(X.data_ptr(), tokens, cols, L.MXQ_F16, out.data_ptr(), 0.0, 1.0, 0, ctas,
                                     ws.data_ptr(), ws.numel(), L.stream())
        assert rc == 0
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    assert torch.allclose(out, ref, rtol=1e-5)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        L.lib().mxq_colsumsq_ex(X.data_ptr(), tokens, cols, L.MXQ_F16, out.data_ptr(), 0.0, 1.0, 0, ctas,
                                ws.data_ptr(), ws.numel(), L.stream())
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"ctas_per_sm={ctas}: {ms * 1e3:.1f} us = {X.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
print("ok")
