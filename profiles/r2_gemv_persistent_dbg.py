"""Where the persistent GEMV chain spends its time: same-shape chains under the profiling modes of
csrc/gemv_chain.cu (MXQ_CHAIN_DBG: 1 = no arithmetic, 2 = no copies, 4 = no activation image)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from profiles.r2_gemv_persistent import graph_time, rand_packed, dev  # noqa: E402


def main():
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for oc, ic in shapes:
        n = 32 if oc * ic < 3e7 else 16
        ps = [rand_packed(oc, ic) for _ in range(n)]
        yy = [torch.empty(oc, device=dev, dtype=torch.float16) for _ in range(n)]
        xs = [torch.randn(ic, device=dev).half() for _ in range(n)]
        for share in (1, 0):
            for dbg in (0, 1, 2, 3, 4, 5, 7):
                os.environ["MXQ_CHAIN_DBG"] = str(dbg)
                c = ops.GemvChain([((xs[0] if share else xs[i]), p, y, -1) for i, (p, y) in enumerate(zip(ps, yy))], validate=False)
                us = graph_time(c.run)
                print(f"{oc}x{ic} share_x={share} dbg={dbg} (noarith={dbg & 1} nocopy={(dbg >> 1) & 1} noimg={(dbg >> 2) & 1}): "
                      f"{us / n:.2f} us per linear", flush=True)
        os.environ.pop("MXQ_CHAIN_DBG", None)
        del ps, yy, xs


if __name__ == "__main__":
    main()
