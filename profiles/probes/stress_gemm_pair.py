"""Stress the CTA-pair GEMM: many launches per shape, every result compared with cuBLAS on the
unpacked weights (exposes cross-CTA publication races)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from mxq_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
bad = 0
for M, OC, IC in ((2048, 4096, 4096), (2048, 11008, 4096), (1024, 4096, 11008), (512, 256, 8192), (4096, 1024, 2048)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    p = ops.pack(W)
    Wd = ops.unpack(p).half()
    x = torch.randn(M, IC, device=dev).half()
    ref = (x @ Wd.T).float()
    scale = float(ref.abs().max())
    worst = 0.0
    for it in range(int(os.environ.get("ITERS", "200"))):
        y = ops.gemm(x, p, validate=False)
        err = float((y.float() - ref).abs().max()) / scale
        worst = max(worst, err)
        if err > 2e-3:
            bad += 1
    print(f"{M}x{OC}x{IC}: worst rel err {worst:.2e}", flush=True)
print("BAD" if bad else "OK", bad)
