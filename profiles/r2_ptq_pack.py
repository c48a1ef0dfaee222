"""fasterquant + pack (mxq_ptq_quant_pack) per Llama-2-7B shape: the fused 16-row-tile kernel vs
round 1's dead-mask + pooled pre-pass + tile kernel chain (MXQ_PTQ_ROUND1=1).  Graph-timed over
rotating weights (> L2); algorithmic bytes = 4.376 B per weight (read fp16, write fp16 + packed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxq_b200 import ops  # noqa: E402
from mxq_b200.prune import LinearQuantJob, packed_nbytes  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6539.9
for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
    nset = 6
    Ws = [(torch.randn(oc, ic, device=dev) * 0.02).half() for _ in range(nset)]
    stat = torch.rand(ic, device=dev) + 0.1
    stat[7] = 0
    jobs = [LinearQuantJob(oc, ic, dev) for _ in range(2)]
    nb = oc * ic * 4 + packed_nbytes(oc, ic)
    for mode in ("fused", "round1"):
        if mode == "round1":
            os.environ["MXQ_PTQ_ROUND1"] = "1"
        else:
            os.environ.pop("MXQ_PTQ_ROUND1", None)

        def run():
            for i, W in enumerate(Ws):
                jobs[i & 1].run(W, stat)
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3 / nset)
        us = sorted(ts)[2]
        print(f"{oc}x{ic} {mode}: {us:.1f} us per linear = {nb / us / 1e3:.0f} GB/s ({nb / us / 1e3 / HBM:.2f} of HBM)", flush=True)
    os.environ.pop("MXQ_PTQ_ROUND1", None)
