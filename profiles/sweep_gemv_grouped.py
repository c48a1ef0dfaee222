"""Grouped decode GEMV (q/k/v: 3 x 4096^2, gate/up: 2 x 11008x4096): planner choice vs forced
(warps, warps-per-row-group, ring stages), CUDA graph over > L2 worth of distinct weights."""
import itertools, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops
from mxq_b200.prune import packed_nbytes
dev = torch.device("cuda:0")
def rand_packed(oc, ic):
    p = {}
    for k, (s, d) in ops.packed_shapes(oc, ic).items():
        p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half() if d == torch.float16 else \
            torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
    return p
def run(oc, ic, n, nsets):
    sets = [[rand_packed(oc, ic) for _ in range(n)] for _ in range(nsets)]
    x = torch.randn(1, ic, device=dev).half()
    ys = [torch.empty(1, oc, device=dev, dtype=torch.float16) for _ in range(n)]
    def bench():
        for s in sets[:2]:
            ops.gemv_grouped(x, s, outs=ys, validate=False)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for s in sets:
                ops.gemv_grouped(x, s, outs=ys, validate=False)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / 5 / nsets * 1e3
    for k in ("MXQ_GEMV_WARPS", "MXQ_GEMV_WPR", "MXQ_GEMV_STAGES"): os.environ.pop(k, None)
    os.environ["MXQ_GEMV_VERBOSE"] = "1"
    ops.gemv_grouped(x, sets[0], outs=ys, validate=False)
    os.environ.pop("MXQ_GEMV_VERBOSE")
    nb = n * packed_nbytes(oc, ic)
    t = bench()
    print(f"{n} x {oc}x{ic} planner: {t:.2f} us = {nb / t / 1e3:.0f} GB/s", flush=True)
    res = []
    for w, wpr, st in itertools.product((8, 10, 12, 14, 16), (1, 2), (1, 2, 3, 4)):
        os.environ.update(MXQ_GEMV_WARPS=str(w), MXQ_GEMV_WPR=str(wpr), MXQ_GEMV_STAGES=str(st))
        try:
            res.append((bench(), w, wpr, st))
        except Exception as e:
            pass
    for t, w, wpr, st in sorted(res)[:6]:
        print(f"   warps {w} wpr {wpr} stages {st}: {t:.2f} us = {nb / t / 1e3:.0f} GB/s")
run(4096, 4096, 3, 20)
run(11008, 4096, 2, 12)
run(4096, 4096, 1, 60)
run(4096, 11008, 1, 24)
run(11008, 4096, 1, 24)
