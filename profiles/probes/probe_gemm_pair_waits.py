"""Cycles the CTA-pair GEMM's MMA issuer (block 0,0) spends waiting for A (TMA) and for B."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from mxq_b200 import ops
dev = torch.device("cuda:0")
h = torch.zeros(64, dtype=torch.int64).pin_memory()
print("pinned", h.is_pinned(), hex(h.data_ptr()))
os.environ["MXQ_GEMM_DBG_PTR"] = str(h.data_ptr())
M, OC, IC = 2048, 4096, 11008
W = (torch.randn(OC, IC, device=dev) * 0.02).half()
x = torch.randn(M, IC, device=dev).half()
p = ops.pack(W)
for name, fn in (("dense", lambda: ops.gemm_dense(x, W)), ("packed", lambda: ops.gemm(x, p, validate=False))):
    for _ in range(3):
        fn(); torch.cuda.synchronize()
        print(f"{name}: wait_a {int(h[61])} wait_b {int(h[62])} total {int(h[63])} cycles over {IC // 64} K blocks", flush=True)
