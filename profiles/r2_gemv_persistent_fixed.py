"""Fixed cost of one persistent-chain launch vs its per-job slope (4096 x 4096 jobs, shared activation)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from profiles.r2_gemv_persistent import graph_time, rand_packed, dev  # noqa: E402


def main():
    oc = ic = 4096
    ps = [rand_packed(oc, ic) for _ in range(64)]
    yy = [torch.empty(oc, device=dev, dtype=torch.float16) for _ in range(64)]
    x = torch.randn(ic, device=dev).half()
    xs = [torch.randn(ic, device=dev).half() for _ in range(64)]
    tiny = rand_packed(32, 256)
    ty = torch.empty(32, device=dev, dtype=torch.float16)
    tx = torch.randn(256, device=dev).half()
    for dbg in (0, 7):
        os.environ["MXQ_CHAIN_DBG"] = str(dbg)
        c = ops.GemvChain([(tx, tiny, ty, -1)], validate=False)
        print(f"dbg={dbg} one 32x256 job: {graph_time(c.run):.2f} us", flush=True)
        for share in (1, 0):
            prev = None
            for n in (1, 2, 4, 8, 16, 32, 64):
                c = ops.GemvChain([((x if share else xs[i]), ps[i], yy[i], -1) for i in range(n)], validate=False)
                us = graph_time(c.run)
                extra = "" if prev is None else f"   slope {(us - prev[1]) / (n - prev[0]):.2f} us/job"
                print(f"dbg={dbg} share_x={share} n={n:2d}: {us:7.2f} us{extra}", flush=True)
                prev = (n, us)
    os.environ.pop("MXQ_CHAIN_DBG", None)
    # an empty graph node for comparison: a trivial kernel launch
    a = torch.zeros(1, device=dev)
    print(f"graph with one tiny torch kernel: {graph_time(lambda: a.add_(1)):.2f} us")


if __name__ == "__main__":
    main()
