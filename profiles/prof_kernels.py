"""Launch each hot kernel a few times on BASELINE shapes so that ncu can capture them.

    python profiles/prof_kernels.py            # plain run (must exit 0 before profiling)
    ncu --set full ... python profiles/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1:] or ["fq", "ste", "act", "ptq", "pack", "gemv", "stats", "gemm"]
REP = int(os.environ.get("MXQ_PROF_REP", "3"))

if "fq" in which:
    for dt in (torch.float32, torch.bfloat16):
        xs = [(torch.randn(4096, 4096, device=dev) * 0.02).to(dt) for _ in range(REP)]
        for x in xs:
            ops.fakequant_fwd(x)
    x = (torch.randn(4096, 11008, device=dev) * 0.02).bfloat16()
    ops.fakequant_fwd(x)
if "act" in which:
    from mxq_b200 import AsymQuantizer, SymQuantizer
    clip = torch.tensor([-2.0, 2.0])
    xa = torch.randn(2, 2048, 4096, device=dev).bfloat16()
    for _ in range(REP):
        SymQuantizer.apply(xa, clip, 8, False)
        AsymQuantizer.apply(xa, clip, 4, False)
if "ste" in which:
    for dt in (torch.float32, torch.bfloat16):
        x = (torch.randn(4096, 4096, device=dev) * 0.02).to(dt)
        g = torch.randn_like(x)
        for _ in range(REP):
            ops.ste_bwd(g, x, -2.0, 2.0)
if "stats" in which:
    X = torch.randn(32 * 2048, 4096, device=dev, dtype=torch.float16)
    for _ in range(REP):
        ops.colsumsq(X)
if "ptq" in which or "pack" in which:
    W = (torch.randn(4096, 4096, device=dev) * 0.02).half()
    W2 = (torch.randn(4096, 11008, device=dev) * 0.02).half()
    stat = torch.ones(4096, device=dev)
    for _ in range(REP):
        if "ptq" in which:
            ops.ptq_quant(W, stat)
        if "pack" in which:
            ops.ptq_quant_pack(W, stat)
    ops.ptq_quant_pack(W2, None)
if "gemv" in which or "gemm" in which:
    def rand_packed(oc, ic):
        p = {}
        for k, (s, d) in ops.packed_shapes(oc, ic).items():
            if d == torch.float16:
                p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
            else:
                p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
        return p
    if "gemv" in which:
        for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
            ps = [rand_packed(oc, ic) for _ in range(REP)]
            x = torch.randn(1, ic, device=dev).half()
            for p in ps:
                ops.gemv(x, p)
    if "gemm" in which:
        try:
            for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
                p = rand_packed(oc, ic)
                x = torch.randn(2048, ic, device=dev).half()
                for _ in range(REP):
                    ops.gemm(x, p)
        except RuntimeError as e:
            print("gemm skipped:", e)
torch.cuda.synchronize()
print("prof_kernels ok")
