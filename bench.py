#!/usr/bin/env python
"""bench.py -- MXQ quantization hot path on B200 (contract in the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ptq] [--impl ours|reference]

Default workload (BASELINE.json configs[1]): the full mxq quantization pass over a random-init
Llama-2-7B (32 decoder layers x 7 linears) with synthetic 128 x 2048-token calibration
activations, decoder layers sharded `layer % world` over the GPUs, no data-path collective.
One step = one pass over all layers: per layer 4 activation statistics (what MXQGPT.add_batch is
used for) + fasterquant + pack of the 7 linears.  `value` = algorithmic GB/s of
the whole job with inputs resident in HBM; `e2e` = the same with HOST (pinned) calibration
tensors and weights copied in and the quantized + packed weights copied out inside the timed
region.  `components` carries the other BASELINE configs measured on rank 0 in the same run
(fake-quant fwd/bwd GB/s, decode GEMV GB/s, prefill dequant-GEMM TFLOP/s).

Under torchrun (N > 1) every rank runs its shard; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, INTER, LAYERS = 4096, 11008, 32       # Llama-2-7B (configuration_llama.py:85-88)
NSAMPLES, SEQLEN = 128, 2048                   # main.py:26,33 ; prune.py:329


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ---------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    tokens = 8192
    cores = os.cpu_count() or 1
    times, nbytes, desc = [], 0, ""
    for i in range(args.warmup + args.steps):
        dt, nbytes, desc = cb.ptq_layer_sample(HIDDEN, INTER, tokens, threads=cores, seed=i)
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    val = nbytes * len(times) / tot / 1e9
    line = {
        "impl": "reference", "metric": "mxq_quant_pass_hbm_GBps", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "mxq PTQ pass (statistics + fasterquant + pack), Llama-2-7B shapes: 32 layers x 7 linears, "
                        "synthetic 128x2048-token calibration per distinct linear input",
            "hidden": HIDDEN, "intermediate": INTER, "layers": LAYERS, "nsamples": NSAMPLES, "seqlen": SEQLEN,
            "sharding": f"layer % {n} (no data-path collective)",
            "schedule": "statistics of layer i+1 on the main stream overlap quantize+pack of layer i on a side stream",
            "l2": "inputs larger than L2: 12.2 GB of calibration activations + 0.4 GB of weights stream per layer"}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ptq", choices=["ptq", "gemm70b", "qat"])
    ap.add_argument("--no-components", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one stream: statistics then quantize, layer by layer")
    ap.add_argument("--layers", type=int, default=LAYERS, help="(debug) fewer layers; default is the named config")
    ap.add_argument("--nsamples", type=int, default=NSAMPLES, help="(debug) fewer calibration samples")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3 breaks the timing rules", file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "gemm70b":
        return run_gemm70b(args)
    if args.workload == "qat":
        return run_qat(args)

    import torch
    import torch.distributed as dist
    from mxq_b200 import _lib
    _lib.lib()                              # fail loudly if the CUDA library is missing
    from mxq_b200 import prune

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tokens = args.nsamples * SEQLEN
    my_layers = [l for l in range(args.layers) if l % world == rank]
    ptq = prune.LlamaLayerPTQ(HIDDEN, INTER, dev, tokens)
    lin = prune.llama_linears(HIDDEN, INTER)

    # ---- synthetic data, resident in HBM --------------------------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(1000)
    calib = {}
    for key, d in (("attn_in", HIDDEN), ("o_in", HIDDEN), ("mlp_in", HIDDEN), ("down_in", INTER)):
        X = torch.randn((tokens, d), generator=g, device=dev, dtype=torch.float16)
        X[:, 7] = 0                          # dead column (mxqgpt.py:401)
        X[:, 100:108] *= 20                  # LLM-like outlier channels
        calib[key] = X
    weights = {}
    for l in my_layers:
        g.manual_seed(l)
        weights[l] = {name: (torch.randn((oc, ic), generator=g, device=dev, dtype=torch.float32) * 0.02).half()
                      for name, (oc, ic, _) in lin.items()}
    stat_b, quant_b = prune.algorithmic_bytes_per_layer(HIDDEN, INTER, tokens)
    job_bytes = args.layers * (stat_b + quant_b)

    # ---- dominant kernel (column sum-of-squares) timed with events on the launching stream --
    n_stat_calls = 4 * len(my_layers) * args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_stat_calls)]
    ev_bytes = []
    cursor = {"i": 0, "on": False}

    def on_stat(key, X2, before):
        if not cursor["on"]:
            return
        a, b = ev[cursor["i"]]
        if before:
            a.record()
        else:
            b.record()
            ev_bytes.append(X2.numel() * X2.element_size())
            cursor["i"] += 1

    def step():
        if args.serial:
            for l in my_layers:
                ptq.run(weights[l], calib, args.nsamples, on_stat=on_stat)
        else:
            # statistics of layer i+1 (HBM-bound) overlap quantize+pack of layer i (issue-bound);
            # 3 = the statistics variant with 16 loads in flight per thread, 2 CTAs resident per SM, 1.5
            # waves: 69.7 ms per pass against 71.4 for the full-occupancy kernel (profiles/r1_sweep_stat_overlap.txt)
            ptq.run_pipelined(((weights[l], calib) for l in my_layers), args.nsamples, on_stat=on_stat,
                              ctas_per_sm=int(os.environ.get("MXQ_STAT_CTAS", "3")))

    for _ in range(args.warmup):
        step()
    clocks = Clocks(local)
    barrier()
    clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cursor["on"] = True
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    barrier()
    cursor["on"] = False
    clk = clocks.stop()
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    ms_step = ms_total / args.steps
    value = job_bytes / (ms_step * 1e-3) / 1e9

    pk = peaks()
    stat_ms = [a.elapsed_time(b) for a, b in ev[:cursor["i"]]]
    dom_bytes = sum(ev_bytes) / max(len(ev_bytes), 1)
    dom_ms = sum(stat_ms) / max(len(stat_ms), 1)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "colsumsq_partial_kernel<__half> (+ its 1-block-wide finalize, same event pair)",
                "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "peak_source": pk["src"] + " (burst copy bandwidth, MEASURED_PEAKS.json)" if pk["src"] == "measured" else "fallback 6.65 TB/s",
                # dram__bytes_read+write per launch from the ncu --set full capture in
                # profiles/r1c_ncu_full_bench_kernels.csv (2.152 GB + 4.9 MB for the 2.147 GB input,
                # 5.77 GB + 7.9 MB for the 5.771 GB one): no re-reads
                "traffic": 1.0025 * dom_bytes, "traffic_source": "profiles/r1c_ncu_full_bench_kernels.csv",
                "note": "timed inside the step, where the quantize+pack kernel of the previous layer shares the SMs with it; alone (--serial) it runs at 0.97",
                "bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                "share_of_step": sum(stat_ms) / max(args.steps, 1) / ms_step if ms_step > 0 else None}

    # ---- end to end: host (pinned) inputs, device->host results, copies inside the timing ---
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(torch, dev, ptq, lin, calib, weights, my_layers, args, job_bytes, barrier, max_over_ranks)
        except Exception as e:  # report, never fake
            e2e = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "error": repr(e)[:300]}

    # ---- other BASELINE configs, rank 0 only -------------------------------------------------
    components = None
    cpu = None
    if rank == 0:
        if not args.no_components:
            del calib
            torch.cuda.empty_cache()
            components = run_components(torch, dev, pk)
        if world == 1:
            from oracle import cpu_baseline as cb
            cores = os.cpu_count() or 1
            dt, nb, desc = cb.ptq_layer_sample(HIDDEN, INTER, 8192, threads=cores)
            cpu = {"value": nb / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc,
                   "seconds": dt}

    if rank == 0:
        line = {
            "metric": "mxq_quant_pass_hbm_GBps", "value": value, "unit": "GB/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world) if (args.layers == LAYERS and args.nsamples == NSAMPLES) else
            dict(workload_config(world), layers=args.layers, nsamples=args.nsamples, note="REDUCED debug config"),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk,
            "gpu_launches": prune.LlamaLayerPTQ.LAUNCHES_PER_LAYER * args.layers * args.steps,
            "job_bytes_per_step": job_bytes, "layers_per_s": args.layers / (ms_step * 1e-3),
            "components": components,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gemm70b(args):
    """BASELINE configs[4]: Llama-2-70B-shape packed dequant-GEMM (M = 2048), output columns sharded
    over the ranks, result all-gathered over NVLink -- GEMM only, GEMM + NCCL all-gather, and the
    fused peer-store epilogue.  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    from mxq_b200 import dist as mdist, ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    M = 2048
    shapes = {"q/o_proj": (8192, 8192), "k/v_proj": (1024, 8192), "gate/up_proj": (28672, 8192), "down_proj": (8192, 28672)}
    out = {}
    tot_flops, tot_ms = 0.0, {"gemm": 0.0, "nccl": 0.0, "p2p": 0.0, "mc": 0.0}

    def rand_packed(oc, ic):
        p = {}
        for k, (s, d) in ops.packed_shapes(oc, ic).items():
            if d == torch.float16:
                p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
            else:
                p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
        return p

    graph_note = {}

    def timed(fn, iters, graph=False, tag=""):
        """Device time per call, max over ranks.  graph=True replays `iters` calls captured in one
        CUDA graph (a shard's GEMM at 8 ranks is ~50-100 us, about what the Python call path costs);
        if the capture is refused the loop is timed eagerly and the line says so."""
        for _ in range(max(3, args.warmup)):
            fn()
        torch.cuda.synchronize()
        g = None
        if graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(iters):
                        fn()
                g.replay()
                torch.cuda.synchronize()
            except Exception as e:
                g = None
                graph_note[tag] = "eager (capture refused: " + repr(e)[:80] + ")"
                torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if g is not None:
            g.replay()
        else:
            for _ in range(iters):
                fn()
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for name, (oc, ic) in shapes.items():
        ocl = oc // world
        p = rand_packed(ocl, ic)
        x = torch.randn(M, ic, device=dev).half()
        ws = ops.gemm_workspace(M, ic, ocl, dev)
        y = torch.empty(M, ocl, device=dev, dtype=torch.float16)
        flops = 2.0 * M * oc * ic
        r = {"flops": flops}
        r["gemm_ms"] = timed(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False), args.steps, graph=True, tag="gemm")
        modes = ["nccl", "p2p", "mc"] if world > 1 else []
        for mode in modes:
            try:
                lin = mdist.ColumnShardedMXQLinear(p, oc, mode=mode)
                lin(x)
                # NCCL mode stays eager (the collective's own launch path is part of it); the fused
                # modes are this repo's kernels + the symmetric-memory barrier kernel
                r[mode + "_ms"] = timed(lambda: lin(x), args.steps, graph=(mode != "nccl"), tag=mode)
            except Exception as e:
                r[mode + "_error"] = repr(e)[:200]
        r["gemm_TFLOPs"] = flops / r["gemm_ms"] / 1e9
        for mode in modes:
            if mode + "_ms" in r:
                r[mode + "_TFLOPs"] = flops / r[mode + "_ms"] / 1e9
                tot_ms[mode] += r[mode + "_ms"]
        tot_ms["gemm"] += r["gemm_ms"]
        tot_flops += flops
        out[name] = r
        del p, x, y
    # exchange reported = the fastest mode that ran on every shape (all three are this repo's path:
    # the GEMM kernel + NCCL all-gather, + fused peer stores, + fused multicast stores)
    complete = [m for m in ("nccl", "p2p", "mc") if world > 1 and all((m + "_ms") in r for r in out.values())]
    best = min(complete, key=lambda m: tot_ms[m]) if complete else "gemm"
    value = tot_flops / tot_ms[best] / 1e9
    if rank == 0:
        line = {"metric": "mxq_dequant_gemm_70b_TFLOPs", "value": value, "unit": "TFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": tot_ms[best], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "Llama-2-70B-shape packed mixed 2/4-bit dequant-GEMM, M=2048, output columns sharded + gathered",
                           "exchange": best, "shapes": {k: list(v) for k, v in shapes.items()}},
                "roofline": {"bound": "tensor", "achieved": tot_flops / tot_ms["gemm"] / 1e9 / world, "peak": pk["tf_burst"],
                             "unit": "TFLOP/s", "frac": tot_flops / tot_ms["gemm"] / 1e9 / world / pk["tf_burst"],
                             "traffic": None, "note": "per-GPU GEMM-only rate vs measured cuBLAS bf16 burst peak"},
                "gemm_only_TFLOPs": tot_flops / tot_ms["gemm"] / 1e9,
                "nccl_TFLOPs": tot_flops / tot_ms["nccl"] / 1e9 if "nccl" in complete else None,
                "p2p_TFLOPs": tot_flops / tot_ms["p2p"] / 1e9 if "p2p" in complete else None,
                "mc_TFLOPs": tot_flops / tot_ms["mc"] / 1e9 if "mc" in complete else None,
                "per_shape": out, "gpu_launches": args.steps * len(shapes),
                "timing": "CUDA graph of `steps` calls per mode (NCCL mode eager)" + ("; " + str(graph_note) if graph_note else "")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_qat(args):
    """BASELINE configs[3]: one LLM-QAT `run_train.sh 2 32 32` optimisation step on a random-init
    Llama-2-7B (bf16, batch 2 x 2048 tokens per GPU, gradient checkpointing, KD against a frozen
    teacher, AdamW), fake-quant through the fused kernels, data parallel with NCCL all-reduce."""
    import torch
    import torch.distributed as dist
    from mxq_b200 import ops, qat
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    layers = args.layers
    cfg = qat.llama_config(layers=layers)
    student, teacher, nq = qat.build_models(cfg, dev)
    model = student
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(student, device_ids=[local], gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(student.parameters(), lr=2e-5, betas=(0.9, 0.95), weight_decay=0.0)
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    B, T = 2, 2048
    ids = [torch.randint(0, 32000, (B, T), generator=g, device=dev) for _ in range(4)]
    h_ids = [t.cpu().pin_memory() for t in ids]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for i in range(max(3, args.warmup)):
        losses.append(qat.qat_step(model, teacher, ids[i % 4], opt))
    clocks = Clocks(local)
    barrier()
    clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        losses.append(qat.qat_step(model, teacher, ids[i % 4], opt))
    t1.record()
    barrier()
    clk = clocks.stop()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # e2e: token ids from pinned host memory in, loss value out, every step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        x = h_ids[i % 4].to(dev, non_blocking=True)
        float(qat.qat_step(model, teacher, x, opt).item())
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    ems = float(ems.item())
    # the fake-quant share: all quantized weights, 2 forwards (checkpoint recompute) + 1 STE backward
    ws = [m.weight.detach() for m in student.modules() if isinstance(m, qat.QuantizeLinear)]
    gs = {w.shape: torch.randn_like(w) for w in ws}
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for w in ws:
        ops.fakequant_fwd(w)
        ops.fakequant_fwd(w)
        ops.ste_bwd(gs[w.shape], w, -2.0, 2.0)
    f1.record()
    torch.cuda.synchronize()
    fq_ms = f0.elapsed_time(f1)
    fq_bytes = sum(w.numel() for w in ws) * 2 * (2 * 2 + 3)
    tokens = B * T * world
    if rank == 0:
        line = {"metric": "qat_step_tokens_per_s", "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "LLM-QAT run_train.sh 2 32 32 step (KD vs frozen teacher, AdamW, gradient checkpointing)",
                           "model": f"Llama-2-7B architecture, {layers} layers, random init", "per_gpu_batch": B, "seq_len": T,
                           "parallelism": f"dp{world}", "quantized_linears": nq},
                "roofline": {"bound": "hbm", "kernel": "fakequant_row_kernel<bf16> x2 + ste_bwd_kernel<bf16> over all quantized weights",
                             "achieved": fq_bytes / (fq_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": fq_bytes / (fq_ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                             "ms_per_step": fq_ms, "share_of_step": fq_ms / ms},
                "e2e": {"value": tokens / (ems * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4},
                "clocks": clk, "loss_first_last": [float(losses[0]), float(losses[-1])],
                "gpu_launches": args.steps * nq * 3}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, dev, ptq, lin, calib, weights, my_layers, args, job_bytes, barrier, max_over_ranks):
    """Same pass through the public API with every input in pinned host memory."""
    h_calib = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in calib.items()}
    for k in calib:
        h_calib[k].copy_(calib[k])
    h_w = {l: {n: torch.empty(w.shape, dtype=w.dtype, pin_memory=True).copy_(w) for n, w in weights[l].items()}
           for l in my_layers}
    d_calib = calib                                   # the resident device tensors double as staging buffers
    d_w = {n: torch.empty((oc, ic), dtype=torch.float16, device=dev) for n, (oc, ic, _) in lin.items()}
    h_out = {}
    for n, (oc, ic, _) in lin.items():
        job = ptq.jobs[(oc, ic)]
        h_out[n] = (torch.empty((oc, ic), dtype=torch.float16, pin_memory=True),
                    {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in job.packed.items()})
    h2d = sum(v.numel() * v.element_size() for v in h_calib.values()) + \
        sum(w.numel() * 2 for w in h_w[my_layers[0]].values()) if my_layers else 0
    d2h = sum(t.numel() * 2 + sum(v.numel() * v.element_size() for v in p.values()) for t, p in h_out.values())

    def sink(name, Wq, packed):
        ho, hp = h_out[name]
        ho.copy_(Wq, non_blocking=True)
        for k in packed:
            hp[k].copy_(packed[k], non_blocking=True)

    # Double-buffered: a copy stream uploads layer i+1 (12.6 GB over PCIe, the floor of this leg) while
    # layer i is computed and its results go back on the other DMA direction.
    cur = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(d_calib, d_w),
            ({k: torch.empty_like(v) for k, v in d_calib.items()}, {n: torch.empty_like(v) for n, v in d_w.items()})]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def step():
        for b in range(2):
            free[b].record(cur)
        for i, l in enumerate(my_layers):
            b = i & 1
            dc, dw = bufs[b]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])         # layer i-2 no longer reads these buffers
                for k in dc:
                    dc[k].copy_(h_calib[k], non_blocking=True)
                for n in dw:
                    dw[n].copy_(h_w[l][n], non_blocking=True)
                ready[b].record(copy_stream)
            cur.wait_event(ready[b])
            ptq.statistics(dc, args.nsamples)
            ptq.quantize(dw, sink)
            free[b].record(cur)
        cur.synchronize()                               # results are on the host

    e_steps = max(1, min(args.steps, 2))
    step()                                            # warm-up (page-locked paths, first touches)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for _ in range(e_steps):
        step()
    t1.record()
    barrier()
    wall = time.perf_counter() - w0
    ms = max_over_ranks(t0.elapsed_time(t1)) / e_steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"value": job_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "steps": e_steps,
            "h2d_bytes_per_step": h2d * len(my_layers) * 1 if my_layers else 0,
            "d2h_bytes_per_step": d2h * len(my_layers), "ranks_reported": "rank 0 bytes; every rank moves the same per layer",
            "wall_s": wall, "api": "mxq_b200.prune.LlamaLayerPTQ.statistics/quantize with pinned host tensors; uploads of layer i+1 overlap layer i and its downloads"}


def run_components(torch, dev, pk):
    """The other BASELINE configs on one GPU: short, device-timed, inputs rotated beyond L2."""
    from mxq_b200 import ops
    out = {}

    def timeit(fn, iters, warm=3):
        """`iters` launches captured in one CUDA graph and replayed: the Python/ctypes call path
        costs about as much as these 15-60 us kernels, and a launch-rate-bound loop would time the
        host.  Median of 3 replays, CUDA events on the replay stream."""
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / iters)
        return sorted(ts)[1]

    # config 0: fake-quant fwd + STE bwd on a Llama-2-7B q_proj, fp32 and bf16; 6 rotating sets
    for dt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        nset = 6
        xs = [(torch.randn(4096, 4096, device=dev) * 0.02).to(dt) for _ in range(nset)]
        gs = [torch.randn(4096, 4096, device=dev).to(dt) for _ in range(nset)]
        esz = xs[0].element_size()
        outs = [torch.empty_like(xs[0]) for _ in range(nset)]
        lib = ops.L.lib()

        def fwd(i):
            k = i % nset
            ops.L.check(lib.mxq_fakequant_fwd(xs[k].data_ptr(), outs[k].data_ptr(), None, 4096, 4096,
                                              ops.L.dtype_enum(xs[k]), 16, 2, None, ops.L.stream()), "fq")

        def bwd(i):
            k = i % nset
            ops.L.check(lib.mxq_ste_bwd(gs[k].data_ptr(), xs[k].data_ptr(), outs[k].data_ptr(), 4096 * 4096,
                                        ops.L.dtype_enum(xs[k]), -2.0, 2.0, ops.L.stream()), "ste")
        mf, mb = timeit(fwd, 30), timeit(bwd, 30)
        nb = 4096 * 4096 * esz
        out[f"fakequant_fwd_{name}"] = {"ms": mf, "GBps": 2 * nb / mf / 1e6, "frac_hbm": 2 * nb / mf / 1e6 / pk["hbm"]}
        # BASELINE configs[0] names "group 128": the same positional recipe over 512-column blocks
        # ({2,2,2 | pooled 4} groups of 128 columns; no reference implementation: oracle-checked only)
        try:
            def fwd128(i):
                k = i % nset
                ops.L.check(lib.mxq_fakequant_fwd(xs[k].data_ptr(), outs[k].data_ptr(), None, 4096, 4096,
                                                  ops.L.dtype_enum(xs[k]), 128, 2, None, ops.L.stream()), "fq128")
            m128 = timeit(fwd128, 30)
            out[f"fakequant_fwd_{name}_g128"] = {"ms": m128, "GBps": 2 * nb / m128 / 1e6, "frac_hbm": 2 * nb / m128 / 1e6 / pk["hbm"]}
        except Exception as e:
            out[f"fakequant_fwd_{name}_g128"] = {"error": repr(e)[:200]}
        out[f"ste_bwd_{name}"] = {"ms": mb, "GBps": 3 * nb / mb / 1e6, "frac_hbm": 3 * nb / mb / 1e6 / pk["hbm"]}
        del xs, gs, outs
    # activation / KV-cache fake quantizers on a QAT-sized bf16 activation [2, 2048, 4096]
    try:
        nset = 6
        xs = [torch.randn(2, 2048, 4096, device=dev).to(torch.bfloat16) for _ in range(nset)]
        outs = [torch.empty_like(xs[0]) for _ in range(nset)]
        nb = xs[0].numel() * 2
        for mode, bits in (("sym", 8), ("asym", 4)):
            nseg, seglen, period, valid = ops.segquant_plan(tuple(xs[0].shape), mode, False)
            ms = timeit(lambda i: ops.L.check(ops.L.lib().mxq_segquant_fwd(
                xs[i % nset].data_ptr(), outs[i % nset].data_ptr(), nseg, seglen, ops.L.MXQ_BF16,
                0 if mode == "sym" else 1, bits, period, valid, None, 0, ops.L.stream()), "segquant"), 30)
            out[f"actquant_{mode}{bits}_bf16"] = {"ms": ms, "GBps": 2 * nb / ms / 1e6, "frac_hbm": 2 * nb / ms / 1e6 / pk["hbm"]}
        del xs, outs
    except Exception as e:
        out["actquant"] = {"error": repr(e)[:200]}
    # config 2: decode GEMV over the 7 linears of 8 layers of packed random-bit weights (> L2), CUDA graph
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    nl = 8
    packs = []
    for _ in range(nl):
        layer = []
        for oc, ic in shapes:
            p = {}
            for k, (s, d) in ops.packed_shapes(oc, ic).items():
                if d == torch.float16:
                    p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
                else:
                    p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
            layer.append(p)
        packs.append(layer)
    xin = {4096: torch.randn(1, 4096, device=dev).half(), 11008: torch.randn(1, 11008, device=dev).half()}
    yout = {4096: torch.empty(1, 4096, device=dev, dtype=torch.float16),
            11008: torch.empty(1, 11008, device=dev, dtype=torch.float16)}
    from mxq_b200.prune import packed_nbytes
    gbytes = nl * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in shapes)

    def gemv_all():
        for layer in packs:
            for (oc, ic), p in zip(shapes, layer):
                ops.gemv(xin[ic], p, out=yout[oc], validate=False)
    try:
        ms = timeit(lambda i: gemv_all(), 10, warm=1)
        out["gemv_decode_b1"] = {"ms_per_8_layers": ms, "GBps": gbytes / ms / 1e6, "frac_hbm": gbytes / ms / 1e6 / pk["hbm"],
                                 "launches": nl * len(shapes), "note": "CUDA graph of 56 GEMVs, 0.6 GB of packed weights"}
    except Exception as e:
        out["gemv_decode_b1"] = {"error": repr(e)[:200]}
    # same 8 layers with the linears that share an input grouped per launch: q/k/v, o, gate/up, down
    def gemv_grouped_all():
        for layer in packs:
            ops.gemv_grouped(xin[4096], layer[0:3], outs=yq, validate=False)
            ops.gemv(xin[4096], layer[3], out=yout[4096], validate=False)
            ops.gemv_grouped(xin[4096], layer[4:6], outs=yg, validate=False)
            ops.gemv(xin[11008], layer[6], out=yout[4096], validate=False)
    try:
        yq = [torch.empty(1, 4096, device=dev, dtype=torch.float16) for _ in range(3)]
        yg = [torch.empty(1, 11008, device=dev, dtype=torch.float16) for _ in range(2)]
        ms = timeit(lambda i: gemv_grouped_all(), 10, warm=1)
        out["gemv_decode_b1_grouped"] = {"ms_per_8_layers": ms, "GBps": gbytes / ms / 1e6, "frac_hbm": gbytes / ms / 1e6 / pk["hbm"],
                                         "launches": nl * 4, "note": "q/k/v and gate/up share their input: one grouped launch each (mxq_gemv_grouped)"}
    except Exception as e:
        out["gemv_decode_b1_grouped"] = {"error": repr(e)[:200]}
    # config 2: prefill dequant-GEMM, M = 2048
    try:
        M = 2048
        res = {}
        for (oc, ic), p in zip(shapes[3:6:2] + shapes[6:], (packs[0][3], packs[0][5], packs[0][6])):
            x = torch.randn(M, ic, device=dev).half()
            y = torch.empty(M, oc, device=dev, dtype=torch.float16)
            ws = ops.gemm_workspace(M, ic, oc, dev)
            ms = timeit(lambda i: ops.gemm(x, p, out=y, workspace=ws, validate=False), 10)
            tf = 2.0 * M * oc * ic / ms / 1e9
            # same-shape dense fp16 comparator: cuBLAS through torch.matmul (library GEMM, no dequant)
            Wd = (torch.randn(oc, ic, device=dev) * 0.02).half()
            msc = timeit(lambda i: torch.matmul(x, Wd.t(), out=y), 10)
            tfc = 2.0 * M * oc * ic / msc / 1e9
            res[f"{oc}x{ic}"] = {"ms": ms, "TFLOPs": tf, "frac_tensor_burst": tf / pk["tf_burst"],
                                 "cublas_fp16_dense_TFLOPs": tfc, "frac_of_cublas": tf / tfc}
            del Wd
        out["gemm_prefill_m2048"] = res
    except Exception as e:
        out["gemm_prefill_m2048"] = {"error": repr(e)[:200]}
    return out


if __name__ == "__main__":
    main()
