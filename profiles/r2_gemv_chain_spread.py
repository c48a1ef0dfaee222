"""Per-CTA entry / exit times (globaltimer) of the persistent GEMV chain on bench.py's 56-linear decode
workload: how far apart the 148 CTAs finish.  Needs the trace build (MXQ_CHAIN_TRACE=1 python -m
mxq_b200.build) and MXQ_CHAIN_DBG=8."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MXQ_CHAIN_DBG"] = "8"
from mxq_b200 import ops, _lib as L  # noqa: E402
from profiles.r2_gemv_persistent import rand_packed, dev  # noqa: E402

NEV = 96


def trace():
    buf = (C.c_longlong * (5 * NEV * 4))()
    lib = L.lib()
    lib.mxq_debug_chain_trace.argtypes = [C.c_void_p]
    assert lib.mxq_debug_chain_trace(C.cast(buf, C.c_void_p)) == 0
    return np.frombuffer(buf, dtype=np.int64).reshape(5, NEV * 4).copy()


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    jobs = []
    for _ in range(layers):
        xs = {4096: torch.randn(4096, device=dev).half(), 11008: torch.randn(11008, device=dev).half()}
        for oc, ic in shapes:
            jobs.append((xs[ic], rand_packed(oc, ic), torch.empty(oc, device=dev, dtype=torch.float16), -1))
    c = ops.GemvChain(jobs, validate=False)
    for rep in range(4):
        c.run()
        torch.cuda.synchronize()
        t = trace()[2][:296].reshape(148, 2)
        t0 = t[:, 0].min()
        start, end = t[:, 0] - t0, t[:, 1] - t0
        q = np.percentile(end, [0, 10, 50, 90, 100])
        print(f"run {rep}: entry spread {start.max()} ns; exit min/p10/p50/p90/max = {q.astype(int).tolist()} ns; "
              f"mean exit {end.mean():.0f}, last - mean = {end.max() - end.mean():.0f} ns")
    # column-chunk visits per CTA under the planner's rotation (mxq_gemv_chain_plan): does the exit time follow them?
    work, nxt = np.zeros(148), 0
    for _, p, y, _ in jobs:
        oc, ic = y.numel(), p["weight"].shape[1] * 16
        tb, tr = divmod(oc // 16, 148)
        work += tb * -(-ic // 4096)
        work[(nxt + np.arange(tr)) % 148] += -(-ic // 4096)
        nxt = (nxt + tr) % 148
    print(f"planned chunk visits per CTA: min {work.min():.0f} mean {work.mean():.1f} max {work.max():.0f}; "
          f"correlation with exit time {np.corrcoef(work, end)[0, 1]:.2f}")
    order = np.argsort(end)
    print("earliest CTAs:", order[:8].tolist(), end[order[:8]].tolist())
    print("latest CTAs:  ", order[-8:].tolist(), end[order[-8:]].tolist())


main()
