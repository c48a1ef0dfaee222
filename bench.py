#!/usr/bin/env python
"""bench.py -- MXQ quantization hot path on B200 (contract in the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload ptq|gemm70b|qat] [--no-components] [--no-e2e] [--no-qat] [--no-gemm70b]

Headline (BASELINE.json configs[1]): the full mxq quantization pass over a random-init Llama-2-7B
(32 decoder layers x 7 linears) with synthetic 128 x 2048-token calibration activations, decoder
layers sharded `layer % world` over the GPUs, no data-path collective.  One step = one pass over
all layers: per layer 4 activation statistics (what MXQGPT.add_batch is used for) + fasterquant +
pack of the 7 linears.  `value` = algorithmic GB/s of the whole job with inputs
resident in HBM; `e2e` = the same job driven through `mxq_b200.prune.nas_quant` (the reference's
`--prune_method mxq` driver): token ids from pinned host memory in, fp16 fake-quant + packed weights
back to the host, the calibration activations produced on the device by the layer forwards.

`components` carries the other BASELINE configs measured in the same run, each with its own
`roofline` (achieved / peak / frac / ncu traffic), `cpu_baseline` and `e2e`:
  configs[0] fake-quant forward + STE backward (fp32, bf16, "group 128")         -- rank 0
  configs[2] decode GEMV: the 56 linears of 8 layers as ONE persistent launch (mxq_gemv_chain; the roofline),
             with decoder dependencies inside the launch, per-linear / grouped launches, the reference's own
             kernel and cuBLAS fp16 on the same box; prefill dequant-GEMM at M = 2048      -- rank 0
  configs[3] one LLM-QAT `2 32 32` step, data parallel (NCCL all-reduce) over all ranks
  configs[4] Llama-2-70B-shape dequant-GEMM, output columns sharded over all ranks, exchanged by
             NCCL all-gather / fused peer stores / fused NVSwitch multicast stores
Under torchrun (N > 1) every rank runs its shard; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, INTER, LAYERS = 4096, 11008, 32       # Llama-2-7B (configuration_llama.py:85-88)
NSAMPLES, SEQLEN = 128, 2048                   # main.py:26,33 ; prune.py:329
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "r2_ncu_traffic.csv")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def peak_source(pk):
    return "MEASURED_PEAKS.json (burst copy bandwidth / cuBLAS bf16 burst)" if pk["src"] == "measured" else \
        "fallback of B200_PROFILING.md (6.65 TB/s, 1590 TFLOP/s)"


# ---------------------------------------------------------------------------------------------
# ncu traffic lookup: dram__bytes_read.sum + dram__bytes_write.sum per launch of the SHIPPED kernel
# instantiations, from the committed `ncu --set full` capture (profiles/r2_ncu_traffic.csv, made by
# profiles/r2_prof_traffic.py); the row whose algorithmic bytes match is used.
# ---------------------------------------------------------------------------------------------
def ncu_traffic(kernel_substr: str, tag: str | None = None):
    """(bytes per launch, source string) or (None, reason)."""
    if not os.path.exists(NCU_TRAFFIC):
        return None, "no committed ncu capture"
    best = None
    with open(NCU_TRAFFIC) as f:
        for row in csv.DictReader(f):
            if kernel_substr in row["kernel"] and (tag is None or row.get("tag") == tag):
                best = row
                break
    if best is None:
        return None, f"no ncu row for {kernel_substr} {tag or ''}"
    return float(best["dram_read_bytes"]) + float(best["dram_write_bytes"]), \
        f"profiles/r2_ncu_traffic.csv: {best['kernel'][:60]} [{best.get('tag', '')}]"


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
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

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
# distributed context
# ---------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=self.dev)
        else:
            self.dist = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


def graph_time(torch, fn, iters, warm=3, reps=3):
    """`iters` launches captured in one CUDA graph and replayed: the Python/ctypes call path costs
    about as much as a 15-60 us kernel, and a launch-rate-bound loop would time the host.  Median
    of `reps` replays, CUDA events on the replay stream.  Returns ms per launch."""
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
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters)
    return sorted(ts)[len(ts) // 2]


def event_time(torch, fn, iters, warm=2):
    """Eager loop timed with CUDA events (host-side work included): ms per call."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# ---------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    tokens = int(os.environ.get("MXQ_BENCH_REF_TOKENS", "8192"))
    frac = float(os.environ.get("MXQ_BENCH_REF_ROWS", "1.0"))
    cores = os.cpu_count() or 1
    times, nbytes, desc = [], 0, ""
    for i in range(args.warmup + args.steps):
        dt, nbytes, desc = cb.ptq_layer_sample(HIDDEN, INTER, tokens, threads=cores, seed=i, row_fraction=frac)
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
# headline: the PTQ pass
# ---------------------------------------------------------------------------------------------
def run_ptq_pass(ctx, args, pk):
    torch = ctx.torch
    from mxq_b200 import prune
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    tokens = args.nsamples * SEQLEN
    my_layers = [l for l in range(args.layers) if l % world == rank]
    ptq = prune.LlamaLayerPTQ(HIDDEN, INTER, dev, tokens)
    lin = prune.llama_linears(HIDDEN, INTER)

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

    # dominant kernel (column sum-of-squares) timed with events on the launching stream
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
            ptq.run_pipelined(((weights[l], calib) for l in my_layers), args.nsamples, on_stat=on_stat,
                              ctas_per_sm=int(os.environ.get("MXQ_STAT_CTAS", "3")))

    for _ in range(args.warmup):
        step()
    clocks = Clocks(ctx.local)
    ctx.barrier()
    clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cursor["on"] = True
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    ctx.barrier()
    cursor["on"] = False
    clk = clocks.stop()
    ms_step = ctx.max_over_ranks(t0.elapsed_time(t1)) / args.steps
    value = job_bytes / (ms_step * 1e-3) / 1e9

    stat_ms = [a.elapsed_time(b) for a, b in ev[:cursor["i"]]]
    dom_bytes = sum(ev_bytes) / max(len(ev_bytes), 1)
    dom_ms = sum(stat_ms) / max(len(stat_ms), 1)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # per-launch DRAM traffic of the shipped instantiation: 3 launches on 2.147 GB + 1 on 5.771 GB per layer
    t_small, src = ncu_traffic("colsumsq_partial_kernel", "stats_4096")
    t_big, _ = ncu_traffic("colsumsq_partial_kernel", "stats_11008")
    traffic = (3 * t_small + t_big) / 4 if (t_small and t_big) else None
    roofline = {"bound": "hbm", "kernel": "colsumsq_partial_kernel<__half,16,2> (+ its 1-block-wide finalize, same event pair)",
                "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "peak_source": peak_source(pk), "traffic": traffic, "traffic_source": src,
                "note": "average over the 4 launches per layer, timed inside the step where the quantize+pack kernel of the "
                        "previous layer shares the SMs with it",
                "bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                "share_of_step": sum(stat_ms) / max(args.steps, 1) / ms_step if ms_step > 0 else None}
    del calib, weights, ptq
    torch.cuda.empty_cache()
    return dict(value=value, ms_step=ms_step, job_bytes=job_bytes, roofline=roofline, clocks=clk,
                launches=prune.LlamaLayerPTQ.LAUNCHES_PER_LAYER * args.layers * args.steps)


# ---------------------------------------------------------------------------------------------
# e2e: the same job through the reference's driver, nas_quant(args, model, tokenizer, dev)
# ---------------------------------------------------------------------------------------------
def run_e2e_nas_quant(ctx, args, job_bytes):
    """Random-init Llama-2-7B layers of this rank (layer % world), token ids in pinned host memory
    -> nas_quant (capture, 2 x batched layer forwards with statistics hooks, fasterquant + pack)
    -> fp16 fake-quant weights + packed tensors copied back to pinned host memory."""
    torch = ctx.torch
    import argparse as ap
    from transformers import LlamaConfig, LlamaForCausalLM
    from mxq_b200 import prune
    dev = ctx.dev
    n_layers = len([l for l in range(args.layers) if l % ctx.world == ctx.rank])
    if n_layers == 0:
        n_layers = 1
    cfg = LlamaConfig(hidden_size=HIDDEN, intermediate_size=INTER, num_hidden_layers=n_layers,
                      num_attention_heads=32, num_key_value_heads=32, vocab_size=32000,
                      max_position_embeddings=SEQLEN, rms_norm_eps=1e-5, use_cache=False)
    torch.manual_seed(ctx.rank)

    def build(nl):
        c = LlamaConfig(**{**cfg.to_dict(), "num_hidden_layers": nl})
        with torch.device(dev):
            m = LlamaForCausalLM(c).half().eval()
        m.seqlen = SEQLEN
        return m

    gtok = torch.Generator().manual_seed(1234 + ctx.rank)
    h_ids = torch.randint(0, 32000, (args.nsamples, SEQLEN), generator=gtok).pin_memory()
    bsz = int(os.environ.get("MXQ_NAS_BATCH", "16"))

    def run(model, nsamples, timers, layer_callback=None):
        ids = h_ids[:nsamples].to(dev, non_blocking=True)                      # H2D: the job's input
        loader = [(ids[i:i + 1], None) for i in range(nsamples)]
        a = ap.Namespace(nsamples=nsamples, seed=0, save=None, pack=True)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):                       # the driver prints per linear (prune.py:407-408)
            prune.nas_quant(a, model, None, dev, dataloader=loader, batch_size=bsz, timers=timers, layer_callback=layer_callback)

    # warm-up on a 1-layer model with few samples (cuBLAS / SDPA heuristics, lazy module loading)
    warm = build(1)
    run(warm, min(args.nsamples, 2 * bsz), None)
    del warm
    torch.cuda.empty_cache()
    model = build(n_layers)
    # pinned staging for the results of ONE layer, reused (the host consumer would write them to disk)
    lin0 = prune.find_layers(model.model.layers[0])
    stage_w = {n: torch.empty(m.weight.shape, dtype=torch.float16).pin_memory() for n, m in lin0.items()}
    from mxq_b200 import ops
    stage_p = {n: {k: torch.empty(s, dtype=d).pin_memory() for k, (s, d) in ops.packed_shapes(*m.weight.shape).items()}
               for n, m in lin0.items()}
    d2h_layer = sum(t.numel() * 2 for t in stage_w.values()) + \
        sum(v.numel() * v.element_size() for p in stage_p.values() for v in p.values())
    timers = {}
    # results leave the device layer by layer on a copy stream, under the next forwards (nas_quant calls back as
    # soon as a layer's weights and packed tensors are final); ONE pinned staging set: the copy of a layer (9 ms)
    # is long over when the next layer (270 ms) reports -- the event wait below makes that a guarantee
    copy_stream = torch.cuda.Stream()
    staged = torch.cuda.Event()
    d2h_marks = []

    def ship(i, layer):
        ready = torch.cuda.Event()
        ready.record()                                   # the layer's fasterquant + pack kernels
        staged.synchronize()                             # the previous layer's copies have left the staging buffers
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for n, m in prune.find_layers(layer).items():
                stage_w[n].copy_(m.weight.data, non_blocking=True)
                for k, v in m.mxq_packed.items():
                    stage_p[n][k].copy_(v, non_blocking=True)
            b_.record()
            staged.record()
            d2h_marks.append((a_, b_))

    ctx.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    run(model, args.nsamples, timers, ship)
    torch.cuda.current_stream().wait_stream(copy_stream)   # the last layer's results
    t1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ctx.barrier()
    d2h_ms = sum(a_.elapsed_time(b_) for a_, b_ in d2h_marks)
    ms = ctx.max_over_ranks(t0.elapsed_time(t1))
    del model
    torch.cuda.empty_cache()
    return {"value": job_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "steps": 1,
            "h2d_bytes_per_step": int(h_ids.numel() * 8), "d2h_bytes_per_step": int(d2h_layer * n_layers),
            "breakdown_ms_rank0": {"layer_forwards_incl_statistics_hooks": timers.get("forward_ms"),
                                   "fasterquant_and_pack": timers.get("quant_ms"), "results_to_host (copy stream, under the forwards)": d2h_ms,
                                   "calibration_capture (embedding forwards)": timers.get("capture_ms"),
                                   "rest (host driver gaps)": ms - timers.get("forward_ms", 0.0) - timers.get("quant_ms", 0.0) - timers.get("capture_ms", 0.0)},
            "wall_s": wall, "layers_this_rank": n_layers, "forward_batch": bsz,
            "api": "mxq_b200.prune.nas_quant(args, model, tokenizer, dev, dataloader=..., batch_size=16) on a random-init "
                   "Llama-2-7B (this rank's layers), args.pack=True; token ids from pinned host memory, fp16 + packed results to pinned host memory",
            "note": "the dense layer forwards (2 per layer over 128x2048 tokens, cuBLAS/SDPA) are inside this number; they are not "
                    "part of `value`, whose inputs are resident calibration activations"}


# ---------------------------------------------------------------------------------------------
# components (rank 0): configs[0] and configs[2]
# ---------------------------------------------------------------------------------------------
def _cpu_leg(fn, unit, cores):
    try:
        dt, work, desc = fn()
        return {"value": work / dt, "unit": unit, "cores": cores, "kind": "port", "sample": desc, "seconds": dt}
    except Exception as e:  # report, never fake
        return {"value": None, "unit": unit, "cores": cores, "kind": "port", "error": repr(e)[:200]}


def run_components(ctx, pk, with_cpu=True):
    """The other single-GPU BASELINE configs: short, device-timed, inputs rotated beyond L2."""
    torch = ctx.torch
    dev = ctx.dev
    from mxq_b200 import MXAsymQuantizer, engine, ops
    from mxq_b200.packed_linear import MXQLinear
    from mxq_b200.prune import packed_nbytes
    from oracle import cpu_baseline as cb
    cores = os.cpu_count() or 1
    out = {}
    lib = ops.L.lib()

    def hbm_roof(nbytes, ms, kernel, tag):
        ach = nbytes / ms / 1e6
        tr, src = ncu_traffic(kernel, tag)
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": tr, "traffic_source": src, "algorithmic_bytes": nbytes, "ms_per_launch": ms}

    # ---- configs[0]: fake-quant fwd + STE bwd on a Llama-2-7B q_proj, fp32 and bf16; 6 rotating sets (> L2)
    for dt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        nset = 6
        xs = [(torch.randn(4096, 4096, device=dev) * 0.02).to(dt) for _ in range(nset)]
        gs = [torch.randn(4096, 4096, device=dev).to(dt) for _ in range(nset)]
        esz = xs[0].element_size()
        outs = [torch.empty_like(xs[0]) for _ in range(nset)]
        nb = 4096 * 4096 * esz
        st = ops.L.stream

        def fwd(i, group=16):
            k = i % nset
            ops.L.check(lib.mxq_fakequant_fwd(xs[k].data_ptr(), outs[k].data_ptr(), None, 4096, 4096,
                                              ops.L.dtype_enum(xs[k]), group, 2, None, st()), "fq")

        def bwd(i):
            k = i % nset
            ops.L.check(lib.mxq_ste_bwd(gs[k].data_ptr(), xs[k].data_ptr(), outs[k].data_ptr(), 4096 * 4096,
                                        ops.L.dtype_enum(xs[k]), -2.0, 2.0, st()), "ste")
        mf, mb = graph_time(torch, fwd, 30), graph_time(torch, bwd, 30)
        m128 = graph_time(torch, lambda i: fwd(i, 128), 30)
        # four weights per launch (QAT: q/k/v/o of a decoder layer through FakeQuantGroup) and the same
        # weight under an importance mask (allocate_group_bits) on the row-resident kernel
        m4 = graph_time(torch, lambda i: ops.fakequant_fwd_multi([xs[(i + j) % nset] for j in range(4)],
                                                                 outs=[outs[(i + j) % nset] for j in range(4)]), 12) / 4
        gb = ops.allocate_group_bits(xs[0].half(), torch.rand(4096, device=dev) + 0.1)
        mmask = graph_time(torch, lambda i: ops.fakequant_fwd_multi([xs[i % nset]], pooled_mask=gb, outs=[outs[i % nset]]), 30)
        comp = {"workload": f"MXAsymQuantizer fwd + STE bwd, 4096x4096 {name} (Llama-2-7B q_proj), group 16 {{2,2,2 | pooled 4}}",
                "fwd": {"ms": mf, "roofline": hbm_roof(2 * nb, mf, "fakequant_row_kernel", f"fq_{name}")},
                "bwd": {"ms": mb, "roofline": hbm_roof(3 * nb, mb, "ste_bwd_kernel", f"ste_{name}")},
                "fwd_group128": {"ms": m128, "roofline": hbm_roof(2 * nb, m128, "fakequant_row_kernel", f"fq128_{name}"),
                                 "note": "BASELINE configs[0] names 'group 128': the same recipe over 512-column blocks "
                                         "(no reference implementation; oracle-checked)"},
                "fwd_4_weights_per_launch": {"ms_per_weight": m4, "GBps": 2 * nb / m4 / 1e6, "frac_hbm": 2 * nb / m4 / 1e6 / pk["hbm"],
                                             "note": "mxq_fakequant_fwd_multi: q/k/v/o of a layer in one launch (FakeQuantGroup in the QAT step)"},
                "fwd_importance_mask": {"ms": mmask, "GBps": 2 * nb / mmask / 1e6, "frac_hbm": 2 * nb / mmask / 1e6 / pk["hbm"],
                                        "note": "allocate_group_bits mask (which group of every 4 is pooled) on the row-resident kernel"},
                "GBps_fwd_plus_bwd": 5 * nb / (mf + mb) / 1e6, "frac_hbm_fwd_plus_bwd": 5 * nb / (mf + mb) / 1e6 / pk["hbm"]}
        # e2e: the autograd function with HOST tensors: W and the upstream gradient come from pinned
        # memory, the fake-quantized weight and the STE gradient go back
        hW = xs[0].cpu().pin_memory()
        hG = gs[0].cpu().pin_memory()
        hY, hGi = torch.empty_like(hW).pin_memory(), torch.empty_like(hW).pin_memory()
        clip = torch.tensor([-2.0, 2.0])

        def e2e_call():
            w = hW.to(dev, non_blocking=True).requires_grad_(True)
            y = MXAsymQuantizer.apply(w, clip, 2, False)
            y.backward(hG.to(dev, non_blocking=True))
            hY.copy_(y.detach(), non_blocking=True)
            hGi.copy_(w.grad, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        me = event_time(torch, e2e_call, 5)
        comp["e2e"] = {"value": 5 * nb / me / 1e6, "unit": "GB/s", "ms": me, "h2d_bytes_per_step": 2 * nb, "d2h_bytes_per_step": 2 * nb,
                       "api": "MXAsymQuantizer.apply(W, clip, 2, False) + .backward(g) with pinned host W, g -> host y, grad"}
        if with_cpu:
            comp["cpu_baseline"] = _cpu_leg(lambda: _scaled(cb.fakequant_sample(512, 4096, name, threads=cores), 1e-9), "GB/s", cores)
        out[f"fakequant_{name}"] = comp
        del xs, gs, outs, hW, hG, hY, hGi
    # activation / KV-cache fake quantizers on a QAT-sized bf16 activation [2, 2048, 4096]
    try:
        nset = 6
        xs = [torch.randn(2, 2048, 4096, device=dev).to(torch.bfloat16) for _ in range(nset)]
        outs = [torch.empty_like(xs[0]) for _ in range(nset)]
        nb = xs[0].numel() * 2
        for mode, bits in (("sym", 8), ("asym", 4)):
            nseg, seglen, period, valid = ops.segquant_plan(tuple(xs[0].shape), mode, False)
            ms = graph_time(torch, lambda i: ops.L.check(lib.mxq_segquant_fwd(
                xs[i % nset].data_ptr(), outs[i % nset].data_ptr(), nseg, seglen, ops.L.MXQ_BF16,
                0 if mode == "sym" else 1, bits, period, valid, None, 0, ops.L.stream()), "segquant"), 30)
            out[f"actquant_{mode}{bits}_bf16"] = {"ms": ms, "GBps": 2 * nb / ms / 1e6, "frac_hbm": 2 * nb / ms / 1e6 / pk["hbm"]}
        del xs, outs
    except Exception as e:
        out["actquant"] = {"error": repr(e)[:200]}

    # ---- configs[2] decode: GEMV over the 7 linears of 8 layers of packed random-bit weights (> L2)
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    nl = 8

    def rand_packed(oc, ic):
        p = {}
        for k, (s, d) in ops.packed_shapes(oc, ic).items():
            if d == torch.float16:
                p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
            else:
                p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
        return p
    packs = [[rand_packed(oc, ic) for oc, ic in shapes] for _ in range(nl)]
    xin = {4096: torch.randn(1, 4096, device=dev).half(), 11008: torch.randn(1, 11008, device=dev).half()}
    yout = {4096: torch.empty(1, 4096, device=dev, dtype=torch.float16),
            11008: torch.empty(1, 11008, device=dev, dtype=torch.float16)}
    yq = [torch.empty(1, 4096, device=dev, dtype=torch.float16) for _ in range(3)]
    yg = [torch.empty(1, 11008, device=dev, dtype=torch.float16) for _ in range(2)]
    gbytes = nl * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in shapes)

    # pdl=True: a decode chain over RESIDENT packed weights (no kernel of the chain writes them), so each
    # GEMV may copy its weights into shared memory while its predecessor still runs
    def gemv_all(_):
        for layer in packs:
            for (oc, ic), p in zip(shapes, layer):
                ops.gemv(xin[ic], p, out=yout[oc], validate=False, pdl=True)

    def gemv_grouped_all(_):
        for layer in packs:
            ops.gemv_grouped(xin[4096], layer[0:3], outs=yq, validate=False, pdl=True)
            ops.gemv(xin[4096], layer[3], out=yout[4096], validate=False, pdl=True)
            ops.gemv_grouped(xin[4096], layer[4:6], outs=yg, validate=False, pdl=True)
            ops.gemv(xin[11008], layer[6], out=yout[4096], validate=False, pdl=True)
    gemv = {"workload": "batch-1 decode GEMV over the 56 packed linears of 8 Llama-2-7B layers (0.6 GB of distinct packed weights); "
                        "headline = ONE persistent launch (mxq_gemv_chain), per-linear launches (CUDA graph, PDL) beside it"}
    try:
        # --- the persistent chain: 4 activation vectors per layer (q/k/v | o | gate/up | down), distinct outputs
        xl = [{"attn": torch.randn(4096, device=dev).half(), "o": torch.randn(4096, device=dev).half(),
               "mlp": torch.randn(4096, device=dev).half(), "down": torch.randn(11008, device=dev).half()} for _ in range(nl)]
        xkey = ["attn", "attn", "attn", "o", "mlp", "mlp", "down"]
        ych = [[torch.empty(oc, device=dev, dtype=torch.float16) for oc, _ in shapes] for _ in range(nl)]
        jobs = [(xl[li][xkey[i]], layer[i], ych[li][i], -1) for li, layer in enumerate(packs) for i in range(len(shapes))]
        chain = ops.GemvChain(jobs, validate=False)
        nrep = 8        # several chain launches per graph: one launch per graph would time the replay overhead (~8 us)
        ms0 = graph_time(torch, lambda _: chain.run(), nrep, warm=1, reps=5)
        # back-to-back decode steps: MXQ_GEMV_CHAIN_PDL lets a launch build its tile lists and prefetch weights while
        # the previous launch's CTAs are still leaving (they leave up to 10 us apart); x is read after it has completed
        ms = graph_time(torch, lambda _: chain.run(pdl=True), nrep, warm=1, reps=5)
        ach = gbytes / ms / 1e6
        tr, src = ncu_traffic("gemv_chain_kernel", "gemv_chain_56linear")
        gemv["roofline"] = {"bound": "hbm", "kernel": "gemv_chain_kernel (persistent: TMA stage ring + IMMA m16n8k32, csrc/gemv_chain.cu)",
                            "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                            "traffic": tr, "traffic_source": (src or "") + " (this workload, one launch; profiles/r2j_ncu_full_gemv_chain56.csv)",
                            "algorithmic_bytes": gbytes, "ms_per_8_layers": ms, "launches": 1,
                            "ms_per_8_layers_without_pdl": ms0, "GBps_without_pdl": gbytes / ms0 / 1e6,
                            "note": "independent jobs: the weight-stream rate of one launch over 56 linears, launches back to back "
                                    "(8 per graph replay) with programmatic dependent launch; the same without it beside"}
        # the same 56 linears as a DEPENDENT chain: q/k/v <- x, o <- q, gate/up <- o, down <- gate, next layer <- down
        dj, prev, xcur = [], -1, xl[0]["attn"]
        for li, layer in enumerate(packs):
            base = len(dj)
            q, k_, v, o, gt, up, dn = ych[li]
            dj += [(xcur, layer[0], q, prev), (xcur, layer[1], k_, prev), (xcur, layer[2], v, prev), (q, layer[3], o, base),
                   (o, layer[4], gt, base + 3), (o, layer[5], up, base + 3), (gt, layer[6], dn, base + 4)]
            prev, xcur = base + 6, dn
        dchain = ops.GemvChain(dj, validate=False)
        msd = graph_time(torch, lambda _: dchain.run(), nrep, warm=1, reps=5)
        gemv["chain_dependent"] = {"ms_per_8_layers": msd, "GBps": gbytes / msd / 1e6, "frac_hbm": gbytes / msd / 1e6 / pk["hbm"],
                                   "note": "every linear waits for its producer inside the launch (32 dependency points): "
                                           "the latency of a real decode step, not a bandwidth figure"}
        del chain, dchain
        ms = graph_time(torch, gemv_all, 1, warm=1, reps=5)
        ach = gbytes / ms / 1e6
        tr, src = ncu_traffic("gemv_mxq_kernel", "gemv_4096x4096")
        gemv["per_linear"] = {"ms_per_8_layers": ms, "launches": nl * len(shapes), "GBps": ach, "frac_hbm": ach / pk["hbm"],
                              "kernel": "gemv_mxq_kernel<1> (TMA-ring kernel, one launch per linear, PDL)",
                              "traffic": tr, "traffic_source": (src or "") + " (per 4096x4096 launch: 6,318,080 algorithmic bytes)"}
        ms = graph_time(torch, gemv_grouped_all, 1, warm=1, reps=5)
        ach = gbytes / ms / 1e6
        gemv["grouped"] = {"ms_per_8_layers": ms, "launches": nl * 4, "GBps": ach, "frac_hbm": ach / pk["hbm"],
                           "note": "q/k/v and gate/up share their input: one grouped launch each (mxq_gemv_grouped)"}
        # the integer-tensor-core kernel (csrc/gemv_mma.cu) on the same chains, for the record
        os.environ["MXQ_GEMV_IMPL"] = "mma"
        try:
            m1 = graph_time(torch, gemv_all, 1, warm=1, reps=5)
            m2 = graph_time(torch, gemv_grouped_all, 1, warm=1, reps=5)
            gemv["imma_kernel"] = {"per_linear_GBps": gbytes / m1 / 1e6, "grouped_GBps": gbytes / m2 / 1e6,
                                   "note": "MXQ_GEMV_IMPL=mma: IMMA m16n8k32 inner products + per-warp cp.async rings, one launch per linear"}
        finally:
            os.environ.pop("MXQ_GEMV_IMPL", None)
    except Exception as e:
        gemv["error"] = repr(e)[:200]
    # same-box comparators on 4096x4096, batch 1 (outside the product path)
    try:
        nb1 = packed_nbytes(4096, 4096) + 2 * 8192
        ps = [layer[i] for layer in packs for i in range(4)]
        x1, y1 = xin[4096], yout[4096]
        ms = graph_time(torch, lambda _: [ops.gemv(x1, p, out=y1, validate=False, pdl=True) for p in ps], 1, warm=1, reps=5) / len(ps)
        cmp = {"ours_us": ms * 1e3, "ours_GBps": nb1 / ms / 1e6}
        Ws = [(torch.randn(4096, 4096, device=dev) * 0.02).half() for _ in range(12)]
        msc = graph_time(torch, lambda _: [torch.matmul(x1, W.t(), out=y1) for W in Ws], 1, warm=1, reps=5) / len(Ws)
        cmp["cublas_fp16_dense_us"] = msc * 1e3
        cmp["cublas_fp16_dense_GBps"] = 4096 * 4096 * 2 / msc / 1e6
        del Ws
        so = os.path.join(ROOT, "oracle", "_ref", "mxq_inference_engine.so")
        if os.path.exists(so):
            import importlib.util
            spec = importlib.util.spec_from_file_location("mxq_inference_engine", so)
            ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref)

            def run_ref():
                for p in ps:
                    ref.gemv_mxq_forward_cuda(x1, p["weight"], p["weight_last"], p["zeros_and_scales"], p["scales_2nd"],
                                              p["zeros_2nd"], p["scales_4b"], p["zeros_4b"], 16)
            # the reference launches on the legacy default stream and allocates its output: not capturable in a
            # graph; timed as an eager loop like cuda_kernel/test_mxq_gemv.py:54-61, but with CUDA events
            s0 = torch.cuda.default_stream()
            with torch.cuda.stream(s0):
                msr = event_time(torch, run_ref, 5) / len(ps)
            cmp["reference_kernel_us"] = msr * 1e3
            cmp["reference_kernel_GBps"] = nb1 / msr / 1e6
            cmp["reference_kernel"] = "gemv_mxq_forward_cuda compiled from the reference's own sources (oracle/_ref), IC = 4096, eager loop"
        else:
            cmp["reference_kernel"] = "oracle/_ref/mxq_inference_engine.so not built"
        gemv["same_box_4096x4096_b1"] = cmp
    except Exception as e:
        gemv["same_box_4096x4096_b1"] = {"error": repr(e)[:200]}
    # e2e: the reference binding's call with a HOST activation vector and a host result
    try:
        p = packs[0][0]
        hx = xin[4096].cpu().pin_memory()
        hy = torch.empty(1, 4096, dtype=torch.float16).pin_memory()

        def e2e_gemv():
            y = engine.gemv_mxq_forward_cuda(hx.to(dev, non_blocking=True), p["weight"], p["weight_last"], p["zeros_and_scales"],
                                             p["scales_2nd"], p["zeros_2nd"], p["scales_4b"], p["zeros_4b"], 16)
            hy.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        me = event_time(torch, e2e_gemv, 50)
        gemv["e2e"] = {"value": (packed_nbytes(4096, 4096) + 16384) / me / 1e6, "unit": "GB/s", "ms": me, "h2d_bytes_per_step": 8192,
                       "d2h_bytes_per_step": 8192, "api": "mxq_inference_engine.gemv_mxq_forward_cuda(x, ...) with a pinned host x and host y; "
                                                           "packed weights resident (model state); one call = launch + 2 copies + sync"}
    except Exception as e:
        gemv["e2e"] = {"value": None, "error": repr(e)[:200]}
    if with_cpu:
        gemv["cpu_baseline"] = _cpu_leg(lambda: _scaled(cb.dequant_matmul_sample(1024, 4096, 1, threads=cores, what="bytes"), 1e-9), "GB/s", cores)
    out["gemv_decode_b1"] = gemv

    # ---- configs[2] prefill: dequant-GEMM, M = 2048
    gemm = {"workload": "packed mixed 2/4-bit dequant-GEMM, M = 2048 tokens, Llama-2-7B linear shapes (tcgen05 cta_group::2, TMEM, TMA)"}
    try:
        M = 2048
        res = {}
        tot_f, tot_ms = 0.0, 0.0
        for (oc, ic), p in zip(shapes[3:6:2] + shapes[6:], (packs[0][3], packs[0][5], packs[0][6])):
            x = torch.randn(M, ic, device=dev).half()
            y = torch.empty(M, oc, device=dev, dtype=torch.float16)
            ws = ops.gemm_workspace(M, ic, oc, dev)
            ms = graph_time(torch, lambda i: ops.gemm(x, p, out=y, workspace=ws, validate=False), 10)
            fl = 2.0 * M * oc * ic
            tf = fl / ms / 1e9
            Wd = (torch.randn(oc, ic, device=dev) * 0.02).half()
            msc = graph_time(torch, lambda i: torch.matmul(x, Wd.t(), out=y), 10)
            res[f"{oc}x{ic}"] = {"ms": ms, "TFLOPs": tf, "frac_tensor_burst": tf / pk["tf_burst"],
                                 "cublas_fp16_dense_TFLOPs": fl / msc / 1e9, "frac_of_cublas": msc / ms}
            tot_f += fl
            tot_ms += ms
            del Wd, x, y
        ach = tot_f / tot_ms / 1e9
        tr, src = ncu_traffic("gemm_mxq_pair_kernel", "gemm_4096x4096")
        gemm["per_shape"] = res
        gemm["roofline"] = {"bound": "tensor", "kernel": "gemm_mxq_pair_kernel (+ gemm_split_reduce_kernel for K-split tail tiles)",
                            "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"],
                            "traffic": tr, "traffic_source": (src or "") + " (4096x4096 launch: 6.3 MB packed + 16.8 MB x + 16.8 MB y algorithmic)",
                            "note": "FLOP-weighted over the three shapes; peak = measured cuBLAS bf16 burst"}
        # e2e: MXQLinear with a HOST activation matrix and a host result
        lin = MXQLinear.from_packed(packs[0][3])
        hx = torch.randn(M, 4096).half().pin_memory()
        hy = torch.empty(M, 4096, dtype=torch.float16).pin_memory()

        def e2e_gemm():
            hy.copy_(lin(hx.to(dev, non_blocking=True)), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        me = event_time(torch, e2e_gemm, 10)
        gemm["e2e"] = {"value": 2.0 * M * 4096 * 4096 / me / 1e9, "unit": "TFLOP/s", "ms": me, "h2d_bytes_per_step": M * 4096 * 2,
                       "d2h_bytes_per_step": M * 4096 * 2, "api": "MXQLinear(4096 -> 4096).forward(x) with pinned host x [2048, 4096] and host y"}
    except Exception as e:
        gemm["error"] = repr(e)[:200]
    if with_cpu:
        gemm["cpu_baseline"] = _cpu_leg(lambda: _scaled(cb.dequant_matmul_sample(1024, 4096, 256, threads=cores, what="flops"), 1e-12), "TFLOP/s", cores)
    out["gemm_prefill_m2048"] = gemm
    del packs
    torch.cuda.empty_cache()
    return out


def _scaled(t, k):
    dt, work, desc = t
    return dt, work * k, desc


# ---------------------------------------------------------------------------------------------
# configs[4]: 70B-shape column-sharded dequant-GEMM + exchange (all ranks)
# ---------------------------------------------------------------------------------------------
def run_gemm70b(ctx, args, pk, iters=10):
    torch = ctx.torch
    from mxq_b200 import dist as mdist, ops
    dev, world = ctx.dev, ctx.world
    M = 2048
    shapes = {"q/o_proj": (8192, 8192), "k/v_proj": (1024, 8192), "gate/up_proj": (28672, 8192), "down_proj": (8192, 28672)}
    out = {}
    MODES = ("nccl", "p2p", "mc", "p2p2", "mc2")
    tot_flops, tot_ms = 0.0, dict.fromkeys(("gemm",) + MODES, 0.0)
    graph_note = {}

    def rand_packed(oc, ic):
        p = {}
        for k, (s, d) in ops.packed_shapes(oc, ic).items():
            if d == torch.float16:
                p[k] = (torch.rand(s, device=dev) * 0.009 + 0.001).half()
            else:
                p[k] = torch.randint(-2 ** 31, 2 ** 31 - 1, s, device=dev, dtype=torch.int64).to(torch.int32)
        return p

    def timed(fn, graph, tag):
        """Device time per call, max over ranks.  graph=True replays `iters` calls captured in one CUDA
        graph (a shard's GEMM at 8 ranks is ~50-100 us, about what the Python call path costs)."""
        for _ in range(3):
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
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if g is not None:
            g.replay()
        else:
            for _ in range(iters):
                fn()
        b.record()
        ctx.barrier()
        return ctx.max_over_ranks(a.elapsed_time(b) / iters)

    clocks = Clocks(ctx.local).start()
    for name, (oc, ic) in shapes.items():
        ocl = oc // world
        p = rand_packed(ocl, ic)
        x = torch.randn(M, ic, device=dev).half()
        ws = ops.gemm_workspace(M, ic, ocl, dev)
        y = torch.empty(M, ocl, device=dev, dtype=torch.float16)
        flops = 2.0 * M * oc * ic
        r = {"flops": flops, "tile_bytes_per_rank": M * ocl * 2}
        r["gemm_us"] = 1e3 * timed(lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False), True, "gemm")
        modes = list(MODES) if world > 1 else []
        y_check = {}
        for mode in modes:
            try:
                lin = mdist.ColumnShardedMXQLinear(p, oc, mode=mode)
                y_check[mode] = lin(x).clone()
                # NCCL mode stays eager (the collective's own launch path is part of it); the fused modes are
                # this repo's kernels + the symmetric-memory barrier kernel
                r[mode + "_us"] = 1e3 * timed(lambda: lin(x), mode != "nccl", mode)
            except Exception as e:
                r[mode + "_error"] = repr(e)[:200]
        # correctness of the exchange on THIS run's data: every mode's gathered [M, N] result on every rank against
        # the NCCL all-gather of the plain GEMM (fused modes are bit-identical to it, phased modes add K slices in fp32)
        if "nccl" in y_check:
            ref = y_check["nccl"].float()
            scale = float(ref.abs().max())
            worst = {}
            for mode, yv in y_check.items():
                if mode != "nccl":
                    worst[mode] = ctx.max_over_ranks(float((yv.float() - ref).abs().max()) / max(scale, 1e-30))
            r["max_rel_diff_vs_nccl"] = worst
            r["verified"] = bool(all(v <= 1e-3 for v in worst.values()))
        del y_check
        for mode in modes:
            if mode + "_us" in r:
                tot_ms[mode] += r[mode + "_us"] * 1e-3
        tot_ms["gemm"] += r["gemm_us"] * 1e-3
        tot_flops += flops
        out[name] = r
        del p, x, y
    clk = clocks.stop()
    complete = [m for m in MODES if world > 1 and all((m + "_us") in r for r in out.values())]
    best = min(complete, key=lambda m: tot_ms[m]) if complete else "gemm"
    tile_bytes = sum(r["tile_bytes_per_rank"] for r in out.values())
    res = {"workload": "Llama-2-70B-shape packed mixed 2/4-bit dequant-GEMM, M = 2048, output columns sharded over the ranks + gathered",
           "n_gpus": world, "best_exchange": best,
           "TFLOPs": tot_flops / tot_ms[best] / 1e9, "gemm_only_TFLOPs": tot_flops / tot_ms["gemm"] / 1e9,
           "nccl_TFLOPs": tot_flops / tot_ms["nccl"] / 1e9 if "nccl" in complete else None,
           "p2p_TFLOPs": tot_flops / tot_ms["p2p"] / 1e9 if "p2p" in complete else None,
           "mc_TFLOPs": tot_flops / tot_ms["mc"] / 1e9 if "mc" in complete else None,
           "p2p_phased_TFLOPs": tot_flops / tot_ms["p2p2"] / 1e9 if "p2p2" in complete else None,
           "mc_phased_TFLOPs": tot_flops / tot_ms["mc2"] / 1e9 if "mc2" in complete else None,
           "modes": "nccl = GEMM + NCCL all-gather; p2p / mc = exchange stores fused into the GEMM epilogue (peer stores / NVSwitch "
                    "multicast); p2p2 / mc2 = phased: groups of tiles computed as K slices, each group's reduce + exchange pass on a "
                    "side stream under the next group's tensor work (mxq_gemm_partials / mxq_gemm_reduce_store)",
           "per_shape": out, "verified": (all(r.get("verified", False) for r in out.values()) if world > 1 else None),
           "nvlink_bytes_per_rank": {"egress_p2p": tile_bytes * (world - 1), "egress_mc": tile_bytes if world > 1 else 0,
                                     "ingress": tile_bytes * (world - 1)},
           "roofline": {"bound": "tensor", "achieved": tot_flops / tot_ms["gemm"] / 1e9 / world, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                        "frac": tot_flops / tot_ms["gemm"] / 1e9 / world / pk["tf_burst"], "traffic": None,
                        "note": "per-GPU GEMM-only rate vs measured cuBLAS bf16 burst peak"},
           "clocks": clk,
           "timing": f"CUDA graph of {iters} calls per mode (NCCL mode eager); device time, max over ranks" +
                     ("; " + str(graph_note) if graph_note else "")}
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------
# configs[3]: one LLM-QAT `run_train.sh 2 32 32` step, data parallel (all ranks)
# ---------------------------------------------------------------------------------------------
def run_qat(ctx, args, pk, steps=3, layers=LAYERS):
    torch = ctx.torch
    from mxq_b200 import ops, qat
    dev, world, local, rank = ctx.dev, ctx.world, ctx.local, ctx.rank
    cfg = qat.llama_config(layers=layers)
    student, teacher, nq = qat.build_models(cfg, dev)
    model = student
    if world > 1:
        # 100 MB buckets: 13.5 GB of bf16 gradients in 25 MB pieces reach 325 GB/s of bus bandwidth on 8 GPUs,
        # larger messages use NVLink / NVLS better and still leave >100 buckets to overlap with the backward pass
        model = torch.nn.parallel.DistributedDataParallel(student, device_ids=[local], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=int(os.environ.get("MXQ_DDP_BUCKET_MB", "100")))
    opt = torch.optim.AdamW(student.parameters(), lr=2e-5, betas=(0.9, 0.95), weight_decay=0.0)
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    B, T = 2, 2048
    ids = [torch.randint(0, 32000, (B, T), generator=g, device=dev) for _ in range(4)]
    h_ids = [t.cpu().pin_memory() for t in ids]

    def timed_steps(fn, n):
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        ctx.barrier()
        return ctx.max_over_ranks(a.elapsed_time(b) / n)

    losses = []
    for i in range(3):
        losses.append(qat.qat_step(model, teacher, ids[i % 4], opt))
    clocks = Clocks(local).start()
    ms = timed_steps(lambda i: losses.append(qat.qat_step(model, teacher, ids[i % 4], opt)), steps)
    clk = clocks.stop()
    # e2e: token ids from pinned host memory in, loss value out, every step
    ems = timed_steps(lambda i: float(qat.qat_step(model, teacher, h_ids[i % 4].to(dev, non_blocking=True), opt).item()), steps)
    res_comm = None
    if world > 1:
        # the gradient exchange on its own: the same bytes in DDP-sized buckets through NCCL, device-timed.  The
        # step hides it under the backward pass when ms_per_step(N) - ms_per_step(1) << allreduce_alone_ms.
        grad_bytes = sum(p.numel() * p.element_size() for p in student.parameters() if p.requires_grad)
        flat = torch.empty(grad_bytes // 2, dtype=torch.bfloat16, device=dev)
        chunks = flat.split(int(os.environ.get("MXQ_DDP_BUCKET_MB", "100")) * 1024 * 1024 // 2)

        def allreduce(_):
            for c in chunks:
                ctx.dist.all_reduce(c)
        allreduce(0)
        ar_ms = timed_steps(allreduce, 3)
        res_comm = {"grad_bytes": grad_bytes, "allreduce_alone_ms": ar_ms,
                    "busbw_GBps": 2 * (world - 1) / world * grad_bytes / (ar_ms * 1e-3) / 1e9,
                    "ring_bound_ms": 2 * (world - 1) / world * grad_bytes / 900e9 * 1e3,
                    "note": "torch DDP (100 MB buckets, gradient_as_bucket_view): NCCL all-reduce launched as buckets fill during the "
                            "backward pass.  allreduce_alone_ms = the same bytes in the same buckets with nothing to hide under; the exposed "
                            "part of it is ms_per_step at N GPUs minus ms_per_step at 1 GPU (both in the driver's SCALE record)"}
        del flat, chunks
        # data-parallel correctness on this run: after the optimizer steps every rank must hold the same weights
        chk = torch.stack([p.detach().float().sum() for p in list(student.parameters())[:16]]).double()
        lo, hi = chk.clone(), chk.clone()
        ctx.dist.all_reduce(lo, op=ctx.dist.ReduceOp.MIN)
        ctx.dist.all_reduce(hi, op=ctx.dist.ReduceOp.MAX)
        res_comm["replicas_in_sync"] = bool(torch.equal(lo, hi))
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
    res = {"workload": "LLM-QAT run_train.sh 2 32 32 step (KD vs frozen teacher, AdamW, gradient checkpointing), random-init Llama-2-7B, bf16",
           "n_gpus": world, "parallelism": f"dp{world}", "per_gpu_batch": B, "seq_len": T, "layers": layers, "quantized_linears": nq,
           "tokens_per_s": tokens / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "scaling": "weak",
           "roofline": {"bound": "hbm", "kernel": "fakequant_row_kernel<bf16> x2 + ste_bwd_kernel<bf16> over all quantized weights",
                        "achieved": fq_bytes / (fq_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": fq_bytes / (fq_ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                        "ms_per_step": fq_ms, "share_of_step": fq_ms / ms},
           "e2e": {"value": tokens / (ems * 1e-3), "unit": "tokens/s", "ms": ems, "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4},
           "allreduce": res_comm, "clocks": clk, "loss_first_last": [float(losses[0]), float(losses[-1])],
           "cpu_baseline": None}
    del model, student, teacher, opt, ws, gs
    torch.cuda.empty_cache()
    return res


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
    ap.add_argument("--no-qat", action="store_true")
    ap.add_argument("--no-gemm70b", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one stream: statistics then quantize, layer by layer")
    ap.add_argument("--layers", type=int, default=LAYERS, help="(debug) fewer layers; default is the named config")
    ap.add_argument("--nsamples", type=int, default=NSAMPLES, help="(debug) fewer calibration samples")
    ap.add_argument("--qat-layers", type=int, default=LAYERS, help="(debug) fewer decoder layers in the QAT step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3 breaks the timing rules", file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)

    from mxq_b200 import _lib
    _lib.lib()                              # fail loudly if the CUDA library is missing
    ctx = Ctx()
    torch = ctx.torch
    pk = peaks()
    rank, world = ctx.rank, ctx.world

    if args.workload == "gemm70b":
        r = run_gemm70b(ctx, args, pk, iters=max(args.steps, 5))
        if rank == 0:
            line = {"metric": "mxq_dequant_gemm_70b_TFLOPs", "value": r["TFLOPs"], "unit": "TFLOP/s", "n_gpus": world,
                    "steps": max(args.steps, 5), "warmup": 3, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": {"workload": r["workload"]},
                    "roofline": r["roofline"], "clocks": r["clocks"], "gpu_launches": 4 * max(args.steps, 5), "detail": r}
            print(json.dumps(line), flush=True)
        return ctx.close()
    if args.workload == "qat":
        r = run_qat(ctx, args, pk, steps=max(args.steps, 3), layers=args.qat_layers)
        if rank == 0:
            line = {"metric": "qat_step_tokens_per_s", "value": r["tokens_per_s"], "unit": "tokens/s", "n_gpus": world,
                    "steps": r["steps"], "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": {"workload": r["workload"]},
                    "roofline": r["roofline"], "e2e": r["e2e"], "clocks": r["clocks"], "gpu_launches": r["steps"] * r["quantized_linears"] * 3,
                    "detail": r}
            print(json.dumps(line), flush=True)
        return ctx.close()

    head = run_ptq_pass(ctx, args, pk)

    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e_nas_quant(ctx, args, head["job_bytes"])
        except Exception as e:  # report, never fake
            e2e = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "error": repr(e)[:300]}
            torch.cuda.empty_cache()

    components = {}
    cpu = None
    if not args.no_components and rank == 0:
        try:
            components.update(run_components(ctx, pk, with_cpu=(world == 1)))
        except Exception as e:
            components["error"] = repr(e)[:300]
            torch.cuda.empty_cache()
    if rank == 0 and world == 1:
        from oracle import cpu_baseline as cb
        cores = os.cpu_count() or 1
        dt, nb, desc = cb.ptq_layer_sample(HIDDEN, INTER, 8192, threads=cores)
        cpu = {"value": nb / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc, "seconds": dt}
    if world > 1:
        ctx.barrier()
    # multi-GPU configs run on EVERY rank (collectives inside); at N = 1 they are the single-GPU versions
    if not args.no_components and not args.no_gemm70b:
        try:
            r = run_gemm70b(ctx, args, pk)
            if rank == 0:
                components["gemm70b_sharded"] = r
        except Exception as e:
            if rank == 0:
                components["gemm70b_sharded"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    if not args.no_components and not args.no_qat:
        try:
            r = run_qat(ctx, args, pk, layers=args.qat_layers)
            if rank == 0:
                components["qat_step"] = r
        except Exception as e:
            if rank == 0:
                components["qat_step"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": "mxq_quant_pass_hbm_GBps", "value": head["value"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world) if (args.layers == LAYERS and args.nsamples == NSAMPLES) else
            dict(workload_config(world), layers=args.layers, nsamples=args.nsamples, note="REDUCED debug config"),
            "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": e2e, "clocks": head["clocks"],
            "gpu_launches": head["launches"],
            "job_bytes_per_step": head["job_bytes"], "layers_per_s": args.layers / (head["ms_step"] * 1e-3),
            "components": components or None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
