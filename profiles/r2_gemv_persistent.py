"""Persistent decode-GEMV chain (csrc/gemv_chain.cu) vs per-linear launches on bench.py's 56-linear chain
(8 Llama-2-7B layers, 0.6 GB of distinct packed weights, batch 1), and on same-shape chains.  Independent
jobs (the weight-stream rate) and a fully dependent chain (y of a linear is x of the next)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from mxq_b200.prune import packed_nbytes  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6539.9


def rand_packed(oc, ic):
    p = {}
    for k, (s, d) in ops.packed_shapes(oc, ic).items():
        if d == torch.float16:
            p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
        else:
            p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
    return p


def graph_time(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e3     # us


def main():
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    nl = 8
    packs = [[rand_packed(oc, ic) for oc, ic in shapes] for _ in range(nl)]
    xin = {4096: torch.randn(4096, device=dev).half(), 11008: torch.randn(11008, device=dev).half()}
    gbytes = nl * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in shapes)

    def report(name, us, nbytes=gbytes):
        print(f"{name:58s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s  {nbytes / us / 1e3 / HBM:5.3f} of HBM", flush=True)

    # per-linear launches (the existing default path, PDL)
    yout = {4096: torch.empty(1, 4096, device=dev, dtype=torch.float16), 11008: torch.empty(1, 11008, device=dev, dtype=torch.float16)}

    def per_linear():
        for layer in packs:
            for (oc, ic), p in zip(shapes, layer):
                ops.gemv(xin[ic].view(1, -1), p, out=yout[oc], validate=False, pdl=True)
    report("per-linear launches, ring kernel, PDL (56 launches)", graph_time(per_linear))

    # one persistent launch, independent jobs (distinct outputs)
    ys = [[torch.empty(oc, device=dev, dtype=torch.float16) for oc, _ in shapes] for _ in range(nl)]
    jobs = [(xin[ic], p, y, -1) for layer, yl in zip(packs, ys) for (oc, ic), p, y in zip(shapes, layer, yl)]
    chain = ops.GemvChain(jobs, validate=False)
    report("persistent chain, 56 independent jobs (1 launch)", graph_time(chain.run))
    for pdl in (False, True):                           # back-to-back launches: MXQ_GEMV_CHAIN_PDL overlaps tail and set-up
        report(f"  4 launches per graph, pdl={pdl} (per launch)", graph_time(lambda: [chain.run(pdl=pdl) for _ in range(4)]) / 4)
    # bench.py's layout: four activation vectors per layer (q/k/v | o | gate/up | down), distinct per layer
    xl = [[torch.randn(ic, device=dev).half() for ic in (4096, 4096, 4096, 11008)] for _ in range(nl)]
    xk = [0, 0, 0, 1, 2, 2, 3]
    jobs4 = [(xl[li][xk[i]], p, y, -1) for li, (layer, yl) in enumerate(zip(packs, ys)) for i, (p, y) in enumerate(zip(layer, yl))]
    chain4 = ops.GemvChain(jobs4, validate=False)
    for pdl in (False, True):
        report(f"  32 activation vectors, 4 launches per graph, pdl={pdl}", graph_time(lambda: [chain4.run(pdl=pdl) for _ in range(4)]) / 4)
    del chain4
    # A/B on the same box: images in two fixed buffers (the plan before the image pool)
    for mode in ("0", "2", "3"):
        os.environ["MXQ_CHAIN_IMGPOOL"] = mode
        c3, c4 = ops.GemvChain(jobs, validate=False), ops.GemvChain(jobs4, validate=False)
        os.environ.pop("MXQ_CHAIN_IMGPOOL")
        for name, c in (("3 vectors per layer", c3), ("32 activation vectors", c4)):
            report(f"  IMGPOOL={mode} (0 = two buffers, n = n images ahead), {name}, pdl=True", graph_time(lambda: [c.run(pdl=True) for _ in range(4)]) / 4)
        del c3, c4
    for k in (1, 2, 4):
        sub = ops.GemvChain(jobs[:7 * k], validate=False)
        if k == 1:
            for pdl in (False, True):
                report(f"  7 jobs, 8 launches per graph, pdl={pdl} (per launch)",
                       graph_time(lambda: [sub.run(pdl=pdl) for _ in range(8)]) / 8, gbytes // nl)
        report(f"persistent chain, {7 * k} jobs ({k} layer(s))", graph_time(sub.run), gbytes * k // nl)

    # a real decoder dependency structure: q/k/v <- x; o <- q; gate/up <- o; down <- gate; next layer <- down
    h = [torch.randn(4096, device=dev).half() * 0.01 for _ in range(2)]
    dep_jobs = []
    prev = -1
    xcur = h[0]
    for li, layer in enumerate(packs):
        q, k, v = (torch.empty(4096, device=dev, dtype=torch.float16) for _ in range(3))
        o = torch.empty(4096, device=dev, dtype=torch.float16)
        gt, up = (torch.empty(11008, device=dev, dtype=torch.float16) for _ in range(2))
        dn = torch.empty(4096, device=dev, dtype=torch.float16)
        base = len(dep_jobs)
        dep_jobs += [(xcur, layer[0], q, prev), (xcur, layer[1], k, prev), (xcur, layer[2], v, prev)]
        dep_jobs += [(q, layer[3], o, base)]            # (attention omitted: o_proj reads q)
        dep_jobs += [(o, layer[4], gt, base + 3), (o, layer[5], up, base + 3)]
        dep_jobs += [(gt, layer[6], dn, base + 4)]
        prev = base + 6
        xcur = dn
    dchain = ops.GemvChain(dep_jobs, validate=False)
    report("persistent chain, 56 jobs, decoder dependencies (1 launch)", graph_time(dchain.run))

    # same-shape chains of 32 independent jobs
    for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
        n = 32 if oc * ic < 3e7 else 16
        ps = [rand_packed(oc, ic) for _ in range(n)]
        yy = [torch.empty(oc, device=dev, dtype=torch.float16) for _ in range(n)]
        c = ops.GemvChain([(xin[ic], p, y, -1) for p, y in zip(ps, yy)], validate=False)
        nb = n * (packed_nbytes(oc, ic) + 2 * (oc + ic))
        us = graph_time(c.run)
        report(f"persistent chain, {n} x {oc}x{ic} independent", us, nb)
        print(f"    = {us / n:.2f} us per linear", flush=True)
        del ps, yy, c


if __name__ == "__main__":
    main()
