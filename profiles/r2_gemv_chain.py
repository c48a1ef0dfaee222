"""Decode GEMV, round 2: IMMA kernel (gemv_mma.cu) vs round 1's ring kernel (MXQ_GEMV_IMPL=ring),
per shape and on bench.py's mixed 56-linear chain, each as a CUDA graph over > L2 of distinct
packed weights; comparators on the same box: the reference's own gemv_mxq kernel (oracle/_ref,
IC = 4096 only) and cuBLAS fp16 (torch.matmul on a dense fp16 weight)."""
import importlib.util
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


def graph_time(fn, reps=5):
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
    which = sys.argv[1:] or ["shapes", "chain", "ref"]
    if "early" in which:
        shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
        nl = 8
        packs = [[rand_packed(oc, ic) for oc, ic in shapes] for _ in range(nl)]
        xin = {4096: torch.randn(1, 4096, device=dev).half(), 11008: torch.randn(1, 11008, device=dev).half()}
        yout = {4096: torch.empty(1, 4096, device=dev, dtype=torch.float16),
                11008: torch.empty(1, 11008, device=dev, dtype=torch.float16)}
        gbytes = nl * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in shapes)
        os.environ["MXQ_GEMV_IMPL"] = "mma"
        for early in (0, 1):
            os.environ["MXQ_GEMV_EARLY"] = str(early)
            for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
                ps = [layer[{(4096, 4096): 0, (11008, 4096): 4, (4096, 11008): 6}[(oc, ic)]] for layer in packs] * 3
                us = graph_time(lambda: [ops.gemv(xin[ic], p, out=yout[oc], validate=False, pdl=True) for p in ps]) / len(ps)
                nb = packed_nbytes(oc, ic) + 2 * (oc + ic)
                print(f"{oc}x{ic} early={early}: {us:.2f} us/gemv = {nb / us / 1e3:.0f} GB/s ({nb / us / 1e3 / HBM:.2f})", flush=True)

            def plain():
                for layer in packs:
                    for (oc, ic), p in zip(shapes, layer):
                        ops.gemv(xin[ic], p, out=yout[oc], validate=False, pdl=True)
            us = graph_time(plain)
            print(f"chain 56 launches early={early}: {us:.1f} us = {gbytes / us / 1e3:.0f} GB/s ({gbytes / us / 1e3 / HBM:.2f} of HBM)", flush=True)
        os.environ.pop("MXQ_GEMV_EARLY")
    if "dbg" in which:
        # profiling modes of the IMMA kernel on same-shape chains: MXQ_GEMV_DBG bit 0 = no arithmetic,
        # bit 2 = no copies, bits 8-10 = units issued before the dependency wait
        for oc, ic in ((4096, 4096), (11008, 4096)):
            nset = max(4, int(300e6 / packed_nbytes(oc, ic)))
            ps = [rand_packed(oc, ic) for _ in range(nset)]
            x = torch.randn(1, ic, device=dev).half()
            y = torch.empty(1, oc, device=dev, dtype=torch.float16)
            os.environ["MXQ_GEMV_IMPL"] = "mma"
            for dbg in (0, 16, 1 + 16, 4 + 16, 5 + 16, 16 + 256, 16 + 512, 16 + 768):
                os.environ["MXQ_GEMV_DBG"] = str(dbg)
                us = graph_time(lambda: [ops.gemv(x, p, out=y, validate=False, pdl=True) for p in ps]) / nset
                print(f"{oc}x{ic} dbg={dbg} (noarith={dbg & 1} nocopy={(dbg >> 2) & 1} pre={(dbg >> 8) & 7}): {us:.2f} us/gemv", flush=True)
            os.environ.pop("MXQ_GEMV_DBG")
            del ps

    if "shapes" in which:
        for oc, ic in ((4096, 4096), (11008, 4096), (4096, 11008)):
            nset = max(4, int(300e6 / packed_nbytes(oc, ic)))
            ps = [rand_packed(oc, ic) for _ in range(nset)]
            for B in (1, 4):
                x = torch.randn(B, ic, device=dev).half()
                y = torch.empty(B, oc, device=dev, dtype=torch.float16)
                for impl in ("mma", "ring"):
                    for pdl in (True, False):
                        os.environ["MXQ_GEMV_IMPL"] = impl
                        us = graph_time(lambda: [ops.gemv(x, p, out=y, validate=False, pdl=pdl) for p in ps]) / nset
                        nb = packed_nbytes(oc, ic) + 2 * B * (oc + ic)
                        print(f"{oc}x{ic} B={B} impl={impl} pdl={int(pdl)}: {us:.2f} us/gemv = {nb / us / 1e3:.0f} GB/s "
                              f"({nb / us / 1e3 / HBM:.2f} of HBM)", flush=True)
            del ps
    if "chain" in which:
        shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
        nl = 8
        packs = [[rand_packed(oc, ic) for oc, ic in shapes] for _ in range(nl)]
        xin = {4096: torch.randn(1, 4096, device=dev).half(), 11008: torch.randn(1, 11008, device=dev).half()}
        yout = {4096: torch.empty(1, 4096, device=dev, dtype=torch.float16),
                11008: torch.empty(1, 11008, device=dev, dtype=torch.float16)}
        yq = [torch.empty(1, 4096, device=dev, dtype=torch.float16) for _ in range(3)]
        yg = [torch.empty(1, 11008, device=dev, dtype=torch.float16) for _ in range(2)]
        gbytes = nl * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in shapes)
        for impl in ("auto", "mma", "ring"):
            if impl == "auto":
                os.environ.pop("MXQ_GEMV_IMPL", None)          # the library's per-shape choice
            else:
                os.environ["MXQ_GEMV_IMPL"] = impl
            for pdl in (True, False):
                def plain():
                    for layer in packs:
                        for (oc, ic), p in zip(shapes, layer):
                            ops.gemv(xin[ic], p, out=yout[oc], validate=False, pdl=pdl)

                def grouped():
                    for layer in packs:
                        ops.gemv_grouped(xin[4096], layer[0:3], outs=yq, validate=False, pdl=pdl)
                        ops.gemv(xin[4096], layer[3], out=yout[4096], validate=False, pdl=pdl)
                        ops.gemv_grouped(xin[4096], layer[4:6], outs=yg, validate=False, pdl=pdl)
                        ops.gemv(xin[11008], layer[6], out=yout[4096], validate=False, pdl=pdl)
                for name, fn in (("56 launches", plain), ("grouped, 32 launches", grouped)):
                    us = graph_time(fn)
                    print(f"chain {name} impl={impl} pdl={int(pdl)}: {us:.1f} us per 8 layers = {gbytes / us / 1e3:.0f} GB/s "
                          f"({gbytes / us / 1e3 / HBM:.2f} of HBM)", flush=True)
        del packs
    if "ref" in which:
        oc = ic = 4096
        nset = 48
        ps = [rand_packed(oc, ic) for _ in range(nset)]
        x = torch.randn(1, ic, device=dev).half()
        nb = packed_nbytes(oc, ic) + 2 * (oc + ic)
        so = os.path.join(ROOT, "oracle", "_ref", "mxq_inference_engine.so")
        if os.path.exists(so):
            spec = importlib.util.spec_from_file_location("mxq_inference_engine", so)
            ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref)

            def run_ref():
                for p in ps:
                    ref.gemv_mxq_forward_cuda(x, p["weight"], p["weight_last"], p["zeros_and_scales"], p["scales_2nd"],
                                              p["zeros_2nd"], p["scales_4b"], p["zeros_4b"], 16)
            # the reference launches on the legacy default stream and allocates its output: not graph-capturable;
            # timed as a plain loop, like cuda_kernel/test_mxq_gemv.py:54-61 (but with CUDA events)
            run_ref()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0 = torch.cuda.default_stream()
            with torch.cuda.stream(s0):
                a.record(s0)
                for _ in range(5):
                    run_ref()
                b.record(s0)
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / 5 / nset
            print(f"reference gemv_mxq_forward_cuda 4096x4096 B=1 (oracle/_ref, eager loop): {us:.2f} us = {nb / us / 1e3:.0f} GB/s packed", flush=True)
        else:
            print("oracle/_ref not built")
        Ws = [(torch.randn(oc, ic, device=dev) * 0.02).half() for _ in range(12)]
        y = torch.empty(1, oc, device=dev, dtype=torch.float16)
        us = graph_time(lambda: [torch.matmul(x, W.t(), out=y) for W in Ws]) / len(Ws)
        print(f"cuBLAS fp16 dense GEMV 4096x4096 (torch.matmul, graph): {us:.2f} us = {oc * ic * 2 / us / 1e3:.0f} GB/s dense "
              f"(same linear as packed bytes: {nb / us / 1e3:.0f} GB/s-equivalent)", flush=True)
        os.environ["MXQ_GEMV_IMPL"] = "mma"
        y1 = torch.empty(1, oc, device=dev, dtype=torch.float16)
        us = graph_time(lambda: [ops.gemv(x, p, out=y1, validate=False, pdl=True) for p in ps]) / nset
        print(f"ours (mma, pdl) 4096x4096 B=1: {us:.2f} us = {nb / us / 1e3:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
