"""ncu target for the per-launch DRAM traffic that bench.py reports next to each roofline: ONE launch of
every shipped kernel instantiation the bench times, in a fixed order (r2_make_traffic_csv.py tags
the captured launches by that order).

    ncu --set full --clock-control none -k regex:'colsumsq_partial|fakequant_row|ste_bwd|gemv_mma|gemv_mxq|gemm_mxq_pair|ptq_tile16' \
        -o gpurun_out/r2_traffic -f python profiles/r2_prof_traffic.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402
from mxq_b200.prune import LinearQuantJob  # noqa: E402

dev = torch.device("cuda:0")
L = ops.L
lib = L.lib()
# the order below is the tag order in r2_make_traffic_csv.py
for cols in (4096, 11008):                                     # stats_4096, stats_11008 (bench: 128 x 2048 tokens)
    X = torch.randn(128 * 2048, cols, device=dev, dtype=torch.float16)
    out = torch.empty(cols, device=dev)
    ws = torch.empty(int(lib.mxq_colsumsq_workspace_bytes(X.shape[0], cols)), dtype=torch.uint8, device=dev)
    L.check(lib.mxq_colsumsq_ex(X.data_ptr(), X.shape[0], cols, L.MXQ_F16, out.data_ptr(), 0.0, 1.0, 0, 3, ws.data_ptr(), ws.numel(), L.stream()), "stats")
    torch.cuda.synchronize()
    del X
for dt in (torch.float32, torch.bfloat16):                    # fq_*, fq128_*, ste_*
    x = (torch.randn(4096, 4096, device=dev) * 0.02).to(dt)
    g = torch.randn(4096, 4096, device=dev).to(dt)
    ops.fakequant_fwd(x)
    ops.fakequant_fwd(x, group=128)
    ops.ste_bwd(g, x, -2.0, 2.0)
    torch.cuda.synchronize()


def rand_packed(oc, ic):
    p = {}
    for k, (s, d) in ops.packed_shapes(oc, ic).items():
        if d == torch.float16:
            p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
        else:
            p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
    return p


for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):   # gemv_*, gemm_*
    p = rand_packed(oc, ic)
    ops.gemv(torch.randn(1, ic, device=dev).half(), p)
    ops.gemm(torch.randn(2048, ic, device=dev).half(), p)
    torch.cuda.synchronize()
for oc, ic in ((4096, 4096), (4096, 11008)):                  # ptq_*
    W = (torch.randn(oc, ic, device=dev) * 0.02).half()
    LinearQuantJob(oc, ic, dev).run(W, torch.ones(ic, device=dev))
    torch.cuda.synchronize()
