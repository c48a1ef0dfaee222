"""CTA-pair GEMM: where does the packed path lose time against the dense-B pipeline?
MXQ_GEMM_DBG bits: 16 relay publish, 32 no dequant arithmetic, 64 no weight prefetch loads,
128 no proxy fence, 256 no operand stores."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from mxq_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
M = 2048
for OC, IC in ((4096, 4096), (4096, 11008), (28672, 8192)):
    W = (torch.randn(OC, IC, device=dev) * 0.02).half()
    x = torch.randn(M, IC, device=dev).half()
    p = ops.pack(W)
    y = torch.empty(M, OC, device=dev, dtype=torch.float16)
    ws = torch.zeros(1024, dtype=torch.uint8, device=dev)
    td = timeit(lambda: ops.gemm_dense(x, W))
    print(f"{OC}x{IC}: dense {td:.1f} us ({td*1e3/(IC/64):.0f} ns/kb/wave-ish)")
    for dbg in (0, 512, 480):
        os.environ["MXQ_GEMM_DBG"] = str(dbg)
        tp = timeit(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
        print(f"   dbg={dbg}: packed {tp:.1f} us", flush=True)
    os.environ.pop("MXQ_GEMM_DBG")
