"""Drop-in for ``nas_quant`` / ``find_layers`` (mxq_quant/lib/prune.py:17-36,326-425), the driver
behind ``python main.py --prune_method mxq``, plus ``quantize_linears``: the same per-layer work
(statistics -> fasterquant -> pack) driven directly by calibration tensors, which is what the
benchmark and the layer-sharded multi-GPU pass use."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .mxqgpt import MXQGPT

dataset = "wikitext2"  # prune.py:13


def find_layers(module, layers=(nn.Linear,), name=''):
    """prune.py:17-36 -- recursively collect layers of the given types."""
    if type(module) in layers or any(isinstance(module, t) for t in layers):
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_layers(child, layers=layers, name=name + '.' + name1 if name != '' else name1))
    return res


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


class _StopForward(Exception):
    """Raised by the statistics hook of the last quantizable linear of a layer in the FIRST forward."""


@torch.no_grad()
def nas_quant(args, model, tokenizer, dev, dataloader=None, batch_size: int = 1, timers: dict | None = None,
              early_exit: bool = True, layer_callback=None):
    """prune.py:326-425: calibration capture (Catcher), per layer: MXQGPT per linear, forward hooks
    feeding add_batch during the layer forwards over all samples, fasterquant (blocksize 16), a
    second forward with the quantized weights, in/out swap.

    Extensions: ``dataloader`` -- iterable of (input_ids[1, seqlen], ...) batches (the reference
    downloads wikitext2 here, prune.py:329); ``batch_size`` -- samples per layer forward (the
    reference runs one at a time, :400-402,416-417; statistics and outputs are identical, the dense
    forwards just stop being launch-bound); ``timers`` -- dict that receives the device time of the
    layer forwards (incl. the statistics hooks) and of fasterquant(+pack) in ms; ``args.pack`` --
    also attach the packed 2/4-bit tensors of every linear as ``module.mxq_packed``; ``early_exit`` --
    the outputs of the first forward of a layer are never read (the reference overwrites them with
    the second forward's, :400-402 vs :416-417), so once the LAST linear of the layer has seen its
    input (its statistics are taken in a pre-hook) the rest of that forward -- for Llama the
    down_proj GEMM, 22 % of the layer's linear FLOPs -- is skipped.  Which linear is last is
    observed on the first batch of every layer class (hook firing order); ``layer_callback(i, layer)``
    -- called right after layer i's linears are quantized (their weights / ``mxq_packed`` are final),
    before its second forward: a caller that streams the results out (device -> host, disk) can overlap
    that with the remaining forwards."""
    print('Starting ...')
    if dataloader is None:
        # the reference downloads wikitext2 here (prune.py:329, lib/data.py: needs the `datasets`
        # package and network access); data loading is outside this package's scope
        raise RuntimeError("nas_quant needs a calibration dataloader: pass dataloader=<iterable of "
                           "(input_ids[1, seqlen], ...) batches>")
    use_cache = model.config.use_cache
    model.config.use_cache = False
    layers = model.model.layers
    device_map = getattr(model, "hf_device_map", {})
    if "model.embed_tokens" in device_map:
        dev = device_map["model.embed_tokens"]
    dtype = next(iter(model.parameters())).dtype
    inps = torch.zeros((args.nsamples, model.seqlen, model.config.hidden_size), dtype=dtype, device=dev)
    cache = {'i': 0, 'kwargs': {}}

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, **kwargs):
            b = inp.shape[0]
            inps[cache['i']:cache['i'] + b] = inp
            cache['i'] += b
            cache['kwargs'] = kwargs
            raise ValueError

    marks = []                      # (phase, start event, end event)

    def timed(phase):
        if timers is None:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks.append((phase, a, b))
        a.record()
        return b

    layers[0] = Catcher(layers[0])
    end_capture = timed("capture")
    bs = max(1, int(batch_size))
    # the reference captures one sample per model call (:350-355); with batch_size > 1 (and a sample count it
    # divides, so that every later layer forward sees the batch shape the captured kwargs were made for) the
    # embedding / rotary / mask prologue runs once per batch
    cap = bs if args.nsamples % bs == 0 else 1
    pend = []
    for batch in dataloader:
        pend.append(batch[0].to(dev))
        if len(pend) < cap:
            continue
        try:
            model(torch.cat(pend, dim=0) if len(pend) > 1 else pend[0])
        except ValueError:
            pass
        pend = []
    for ids in pend:                                   # a dataloader shorter than announced: one at a time
        try:
            model(ids)
        except ValueError:
            pass
    layers[0] = layers[0].module
    if end_capture is not None:
        end_capture.record()
    # (the reference calls torch.cuda.empty_cache() here and after the last layer, prune.py:364,423 -- a
    # device-wide synchronisation plus cudaFree of every cached block, needed for its 64-460 MB Hessians)
    outs = torch.zeros_like(inps)
    kwargs = {k: v for k, v in cache['kwargs'].items() if k in ("attention_mask", "position_ids", "position_embeddings")}
    print('Ready.')
    def forward_all(layer):
        end = timed("forward")
        for j in range(0, args.nsamples, bs):
            try:
                out = layer(inps[j:j + bs], **kwargs)
            except _StopForward:                       # first forward only: its outputs are never read
                continue
            outs[j:j + bs] = out[0] if isinstance(out, tuple) else out
        if end is not None:
            end.record()

    last_linear = {}                # layer class -> name of the linear whose hook fires last
    # args.pack: the packed tensors of all linears stay alive (module.mxq_packed); carving them from ONE arena
    # per device replaces ~50 cudaMalloc calls per decoder layer (1 ms each) by one
    arenas = {}
    if getattr(args, "pack", False):
        need = {}
        for i in range(len(layers)):
            for name, lin in find_layers(layers[i]).items():
                d = lin.weight.device
                n = sum(-(-_numel(sh) * dt.itemsize // 256) * 256 for sh, dt in ops.packed_shapes(*lin.weight.shape).values())
                need[d] = need.get(d, 0) + n
        arenas = {d: [torch.empty(n, dtype=torch.uint8, device=d), 0] for d, n in need.items()}

    def carve(lin):
        arena = arenas[lin.weight.device]
        out = {}
        for k, (sh, dt) in ops.packed_shapes(*lin.weight.shape).items():
            n = _numel(sh) * dt.itemsize
            out[k] = arena[0][arena[1]:arena[1] + n].view(dt).view(sh)
            arena[1] += -(-n // 256) * 256
        return out

    for i in range(len(layers)):
        layer = layers[i]
        if f"model.layers.{i}" in device_map:
            dev = device_map[f"model.layers.{i}"]
            kwargs = {k: (tuple(t.to(dev) for t in v) if isinstance(v, tuple) else v.to(dev) if torch.is_tensor(v) else v)
                      for k, v in kwargs.items()}
            inps, outs = inps.to(dev), outs.to(dev)
        subset = find_layers(layer)
        gpts = {name: MXQGPT(subset[name]) for name in subset}

        order = []

        def add_batch(name):
            def tmp(_, inp, out):
                order.append(name)
                gpts[name].add_batch(inp[0].data, out.data)
            return tmp

        def add_batch_and_stop(name):
            def tmp(_, inp):
                gpts[name].add_batch(inp[0].data)
                raise _StopForward
            return tmp

        last = last_linear.get(type(layer)) if early_exit else None
        if last is not None and last not in gpts:
            last = None
        handles = [subset[name].register_forward_hook(add_batch(name)) for name in gpts if name != last]
        if last is not None:
            handles.append(subset[last].register_forward_pre_hook(add_batch_and_stop(last)))
        forward_all(layer)
        for h in handles:
            h.remove()
        if early_exit and last is None and len(order) >= len(gpts) and len(set(order[-len(gpts):])) == len(gpts):
            last_linear[type(layer)] = order[-1]       # every linear fired once per batch: the last one is known
        end = timed("quant")
        for name in gpts:
            print(i, name)
            print('Pruning ...')
            if arenas:
                gpts[name].packed = carve(subset[name])
            gpts[name].fasterquant(percdamp=0.01, blocksize=16,
                                   pack=bool(getattr(args, "pack", False)))
            if getattr(args, "pack", False):
                subset[name].mxq_packed = gpts[name].packed
            gpts[name].free()
        if end is not None:
            end.record()
        if layer_callback is not None:
            layer_callback(i, layer)
        forward_all(layer)
        layers[i] = layer
        inps, outs = outs, inps

    model.config.use_cache = use_cache
    if timers is not None:
        torch.cuda.synchronize()
        for phase, a, b in marks:
            timers[phase + "_ms"] = timers.get(phase + "_ms", 0.0) + a.elapsed_time(b)


class LinearQuantJob:
    """Pre-allocated output buffers for quantizing linears of one shape [OC, IC] repeatedly (no
    allocation inside a timed region)."""

    def __init__(self, OC: int, IC: int, device):
        self.OC, self.IC = OC, IC
        self.Wq = torch.empty((OC, IC), dtype=torch.float16, device=device)
        self.packed = ops.alloc_packed(OC, IC, device)
        lib = ops.L.lib()
        n = max(lib.mxq_ptq_workspace_bytes(OC, IC), lib.mxq_pack_workspace_bytes(OC, IC))
        self.ws = torch.empty(int(n), dtype=torch.uint8, device=device)

    def run(self, W: torch.Tensor, colstat: torch.Tensor):
        """fasterquant + pack of one linear into this job's buffers (W is left untouched)."""
        ops.ptq_quant_pack(W, colstat, out=self.Wq, packed=self.packed, workspace=self.ws)
        return self.Wq, self.packed


# Llama decoder layer: linear name -> (out dim, in dim, which captured input feeds it)
def llama_linears(hidden: int, inter: int, kv_hidden: int | None = None):
    kv = hidden if kv_hidden is None else kv_hidden
    return {
        "self_attn.q_proj": (hidden, hidden, "attn_in"), "self_attn.k_proj": (kv, hidden, "attn_in"),
        "self_attn.v_proj": (kv, hidden, "attn_in"), "self_attn.o_proj": (hidden, hidden, "o_in"),
        "mlp.gate_proj": (inter, hidden, "mlp_in"), "mlp.up_proj": (inter, hidden, "mlp_in"),
        "mlp.down_proj": (hidden, inter, "down_in"),
    }


class LlamaLayerPTQ:
    """The mxq work of one decoder layer given its captured linear inputs: 4 activation statistics
    (what the 7 MXQGPT.add_batch hooks compute, prune.py:389-402) + fasterquant + pack of the 7
    linears (prune.py:404-414).  Kernels launched per layer: 4 x 2 (statistics: partial sums +
    finalize) + 7 x 1 (the fused 16-row-tile fasterquant + pack kernel)."""

    LAUNCHES_PER_LAYER = 4 * 2 + 7

    def __init__(self, hidden: int, inter: int, device, max_tokens: int, kv_hidden=None):
        self.linears = llama_linears(hidden, inter, kv_hidden)
        self.device = device
        self.jobs = {}
        for name, (oc, ic, _) in self.linears.items():
            if (oc, ic) not in self.jobs:
                self.jobs[(oc, ic)] = LinearQuantJob(oc, ic, device)
        self.stats = {k: torch.empty(d, dtype=torch.float32, device=device)
                      for k, d in (("attn_in", hidden), ("o_in", hidden), ("mlp_in", hidden), ("down_in", inter))}
        n = ops.L.lib().mxq_colsumsq_workspace_bytes(max_tokens, hidden)
        n = max(n, ops.L.lib().mxq_colsumsq_workspace_bytes(max_tokens, inter))
        self.stat_ws = torch.empty(int(n), dtype=torch.uint8, device=device)

    def statistics(self, calib: dict, nsamples: int, on_stat=None, buf: int = 0, ctas_per_sm: int = 8):
        stats = self.stats if buf == 0 else self._stats_b()
        for key, X in calib.items():
            X2 = X.reshape(-1, X.shape[-1])
            if on_stat is not None:
                on_stat(key, X2, True)
            with ops.L.on(X2) as st:
                rc = ops.L.lib().mxq_colsumsq_ex(ops.L.ptr(X2), X2.shape[0], X2.shape[1], ops.L.dtype_enum(X2),
                                                 ops.L.ptr(stats[key]), 0.0, 2.0 / nsamples, 0, ctas_per_sm,
                                                 ops.L.ptr(self.stat_ws), self.stat_ws.numel(), st)
            ops.L.check(rc, "mxq_colsumsq")
            if on_stat is not None:
                on_stat(key, X2, False)

    def _stats_b(self):
        if getattr(self, "_stats2", None) is None:
            self._stats2 = {k: torch.empty_like(v) for k, v in self.stats.items()}
        return self._stats2

    def quantize(self, weights: dict, sink=None, buf: int = 0):
        stats = self.stats if buf == 0 else self._stats_b()
        for name, (oc, ic, key) in self.linears.items():
            Wq, packed = self.jobs[(oc, ic)].run(weights[name], stats[key])
            if sink is not None:
                sink(name, Wq, packed)

    def run(self, weights: dict, calib: dict, nsamples: int, sink=None, on_stat=None):
        self.statistics(calib, nsamples, on_stat)
        self.quantize(weights, sink)

    def run_pipelined(self, layers, nsamples: int, sink=None, on_stat=None, ctas_per_sm: int = 3):
        """Several decoder layers back to back: `layers` = iterable of (weights, calib).  The
        HBM-bound statistics of layer i+1 run on the current stream while the issue-bound
        quantize+pack of layer i runs on a high-priority side stream and takes the SM slots the
        statistics CTAs free up (`ctas_per_sm` shapes the statistics grid, see mxq_colsumsq_ex; measured
        on B200: 8 = the full-occupancy kernel, one wave, 71.4 ms / pass; 16 = 71.3; 3 = the variant with
        16 loads in flight per thread, two CTAs resident per SM, 69.7 ms; 0 = shared-memory ring kernel
        92 ms; serial 78.4); statistics are
        double-buffered, the quantizer's output
        buffers are reused layer after layer exactly as in run() (a sink must consume them on the
        side stream)."""
        cur = torch.cuda.current_stream()
        if getattr(self, "_qstream", None) is None:
            self._qstream = torch.cuda.Stream(device=self.device, priority=-1)   # high priority
            self._ev_stat = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_quant = [torch.cuda.Event(), torch.cuda.Event()]
        qs = self._qstream
        qs.wait_stream(cur)
        n = 0
        for i, (weights, calib) in enumerate(layers):
            b = i & 1
            if i >= 2:
                cur.wait_event(self._ev_quant[b])          # statistics buffer b is free again
            self.statistics(calib, nsamples, on_stat, buf=b, ctas_per_sm=ctas_per_sm)
            self._ev_stat[b].record(cur)
            with torch.cuda.stream(qs):
                qs.wait_event(self._ev_stat[b])
                self.quantize(weights, sink, buf=b)
                self._ev_quant[b].record(qs)
            n += 1
        cur.wait_stream(qs)
        return n


def algorithmic_bytes_per_layer(hidden: int, inter: int, tokens: int, kv_hidden=None):
    """statistics read each distinct linear input once (2 B/elt); quantize+pack read
    fp16 W, write fp16 Wq and the 3.0-bit packed tensors (4 + 0.3756 B per weight)."""
    lin = llama_linears(hidden, inter, kv_hidden)
    stats = tokens * (3 * hidden + inter) * 2
    weights = sum(oc * ic for oc, ic, _ in lin.values())
    packed = sum(packed_nbytes(oc, ic) for oc, ic, _ in lin.values())
    return stats, weights * 4 + packed


def packed_nbytes(OC: int, IC: int) -> int:
    return sum(int(torch.Size(s).numel()) * (2 if d == torch.float16 else 4)
               for s, d in ops.packed_shapes(OC, IC).values())
