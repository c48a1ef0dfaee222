"""Soak test of the persistent decode chain: random job lists (shapes incl. multi-chunk rows and fewer tiles
than CTAs, random dependency graphs, shared activations), every list run three times (eager, eager, CUDA
graph) and compared bit for bit across the runs and against per-linear launches within the fp32 re-ordering
tolerance."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from profiles.r2_gemv_persistent import rand_packed, dev  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    widths = [256, 512, 1024, 4096, 8192, 11008]
    pool = {}
    t_start = time.time()
    worst = 0.0
    for it in range(iters):
        n = int(rng.integers(1, 41))
        shapes, deps = [], []
        for j in range(n):
            d = -1
            if j > 0 and rng.random() < 0.6:
                cand = [i for i in range(j) if shapes[i][0] in widths]
                if cand:
                    d = int(rng.choice(cand))
            ic = shapes[d][0] if d >= 0 else int(rng.choice(widths))
            oc = int(rng.choice(widths + [32, 96, 160, 2048]))
            shapes.append((oc, ic))
            deps.append(d)
        packs = []
        for s in shapes:
            if s not in pool:
                p = rand_packed(*s)
                for k in ("scales_2nd", "scales_4b"):
                    p[k] = (p[k].float() * 0.1).half()
                pool[s] = p
            packs.append(pool[s])
        fresh = {w: (torch.randn(w, device=dev) * 0.5).half() for w in widths}
        ys = [torch.zeros(oc, dtype=torch.float16, device=dev) for oc, _ in shapes]
        jobs = [((ys[deps[j]] if deps[j] >= 0 else fresh[shapes[j][1]]), packs[j], ys[j], deps[j]) for j in range(n)]
        chain = ops.GemvChain(jobs, validate=False)
        outs = []
        for rep in range(2):
            for y in ys:
                y.fill_(float("nan"))
            chain.run()
            torch.cuda.synchronize()
            outs.append([y.clone() for y in ys])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                     # two launches back to back, the second admitted early
            chain.run(pdl=True)                       # (MXQ_GEMV_CHAIN_PDL; ignored for chains with dependencies)
            chain.run(pdl=True)
        for y in ys:
            y.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        outs.append([y.clone() for y in ys])
        for j in range(n):
            a, b, c = outs[0][j], outs[1][j], outs[2][j]
            assert torch.isfinite(a).all(), (it, j, shapes[j], deps[j])
            assert torch.equal(a.view(torch.int16), b.view(torch.int16)) and torch.equal(a.view(torch.int16), c.view(torch.int16)), \
                f"iteration {it} job {j}: runs differ"
            xin = outs[0][deps[j]] if deps[j] >= 0 else fresh[shapes[j][1]]
            ref = ops.gemv(xin.view(1, -1), packs[j], validate=False)[0].float()
            # both kernels add exact integer group sums in fp32 (different orders) and round to fp16: a few fp16
            # ulps of the largest output, or of the subnormal spacing when a deep chain has shrunk the values
            diff = float((a.float() - ref).abs().max())
            scale = float(ref.abs().max())
            err = diff / max(scale, 1e-6)
            worst = max(worst, err)
            assert diff <= 2e-3 * scale + 3 * 2.0 ** -24, f"iteration {it} job {j} {shapes[j]} dep {deps[j]}: {diff} vs {scale}"
        if it % 25 == 0:
            print(f"iteration {it}: {n} jobs ok, worst rel diff vs per-linear launches so far {worst:.2e}, {time.time() - t_start:.0f} s", flush=True)
    print(f"{iters} random chains ok; worst rel diff vs per-linear launches {worst:.2e}")


if __name__ == "__main__":
    main()
