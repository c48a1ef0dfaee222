"""AWQ uniform 4-bit decode GEMV (gemv_forward_cuda): time per shape in a CUDA graph over
> L2 worth of distinct weights."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops
dev = torch.device("cuda:0")
G = 128
for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
    zw = -(-(ic // G) // 8)
    nb = oc * ic // 2 + oc * zw * 4 + oc * zw * 8 * 2 + 2 * (oc + ic)
    nset = max(4, int(400e6 / nb))
    sets = [(torch.randint(-2 ** 31, 2 ** 31 - 1, (oc, ic // 8), device=dev, dtype=torch.int64).to(torch.int32),
             (torch.rand(oc, zw * 8, device=dev) * 0.009 + 0.001).half(),
             torch.randint(-2 ** 31, 2 ** 31 - 1, (oc, zw), device=dev, dtype=torch.int64).to(torch.int32)) for _ in range(nset)]
    x = torch.randn(1, ic, device=dev).half()
    for k, s, z in sets[:2]:
        ops.awq_gemv(x, k, s, z, G)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k, s, z in sets:
            ops.awq_gemv(x, k, s, z, G)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 5 / nset * 1e3
    print(f"awq g128 {oc}x{ic} B=1: {us:.2f} us/gemv = {nb / us / 1e3:.0f} GB/s")
