"""clock64 trace of CTA 1 of the persistent GEMV chain (MXQ_CHAIN_DBG & 8): where a compute warp, the
producer, the reducer and a builder warp spend their cycles per stage."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops, _lib as L  # noqa: E402
from profiles.r2_gemv_persistent import rand_packed, dev  # noqa: E402

NEV = 96


def trace():
    buf = (C.c_longlong * (5 * NEV * 4))()
    lib = L.lib()
    lib.mxq_debug_chain_trace.argtypes = [C.c_void_p]
    assert lib.mxq_debug_chain_trace(C.cast(buf, C.c_void_p)) == 0
    return np.frombuffer(buf, dtype=np.int64).reshape(5, NEV, 4).copy()


def main():
    oc, ic = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
    n = 24
    ps = [rand_packed(oc, ic) for _ in range(n)]
    yy = [torch.empty(oc, device=dev, dtype=torch.float16) for _ in range(n)]
    xs = [torch.randn(ic, device=dev).half() for _ in range(n)]
    dep = os.environ.get("TRACE_DEP") == "1"       # job i reads the output of job i - 1 (square shapes only)
    for dbg in ([int(a) for a in sys.argv[3:]] or (8, 9, 10, 15)):
        os.environ["MXQ_CHAIN_DBG"] = str(dbg)
        c = ops.GemvChain([((yy[i - 1] if dep and i else xs[i]), p, y, (i - 1 if dep else -1)) for i, (p, y) in enumerate(zip(ps, yy))], validate=False)
        for _ in range(3):
            c.run()
        torch.cuda.synchronize()
        t = trace()
        t0 = t[1, 0, 0]
        print(f"==== {oc}x{ic} dbg={dbg}: cycles relative to the producer's first stamp")
        print("compute warp 0: [wait full begin, got full, arrived empty, after flush]; d = per-event duration")
        for e in range(40):
            r = t[0, e] - t0
            print(f"  c{e:02d} {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d}   wait {r[1]-r[0]:5d} work {r[2]-r[1]:5d} flush {r[3]-r[2]:5d}")
        print("producer: [wait empty begin, got empty, (copies begin), done]")
        for e in range(40):
            r = t[1, e] - t0
            print(f"  p{e:02d} {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d}   wait {r[1]-r[0]:5d} issue {max(r[3], r[2])-r[1]:5d}")
        print("reducer: [wait begin, got, done]")
        for e in range(24):
            r = t[2, e] - t0
            print(f"  r{e:02d} {r[0]:7d} {r[1]:7d} {r[2]:7d}   wait {r[1]-r[0]:5d} work {r[2]-r[1]:5d}")
        print("compute warp 0 per job: [loop top, after share, after image handshake, before tile loop]")
        for e in range(12):
            r = t[4, e] - t0
            print(f"  j{e:02d} {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d}")
        print("builder: [wait begin, got, done]")
        for e in range(16):
            r = t[3, e] - t0
            print(f"  b{e:02d} {r[0]:7d} {r[1]:7d} {r[2]:7d}   wait {r[1]-r[0]:5d} work {r[2]-r[1]:5d}")


if __name__ == "__main__":
    main()
