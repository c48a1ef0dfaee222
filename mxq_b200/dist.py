"""Multi-GPU plumbing for the sharded paths.  One process per GPU; torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) carries the only exchange step.

* PTQ pass            : decoder layers are independent -> `layer % world`, no data-path collective.
* dequant-GEMM (70B)  : output columns (= weight rows with all their per-row / 4-row / 8-row
                        metadata) are sharded, x is replicated, the [M, N/W] tiles are exchanged:
                        either NCCL all-gather, or the GEMM epilogue stores its tile straight into
                        every peer's output buffer over NVLink (symmetric memory, `mode="p2p"`),
                        or into all of them at once through the NVSwitch multicast mapping of
                        that buffer (`mode="mc"`: one multimem.st, egress = the tile once).
                        `mode="p2p2"` / `"mc2"`: the same stores, but PHASED -- the shard's weight rows
                        are cut into groups of N tiles, every group is computed as K slices that fill the
                        machine (mxq_gemm_partials) and its reduce + exchange pass (mxq_gemm_reduce_store)
                        runs on a side stream under the next group's tensor work.  At 8 ranks a 70B shard
                        is a single partly filled wave, so without phasing the exchange can only start
                        when all of the GEMM is done.
* QAT                 : data parallel, NCCL gradient all-reduce (torch DDP around QuantizeLinear).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ROW_ALIGN = 16   # rows per shard must keep 4-row second-order groups, 8-row zeros_4b words and
                 # 16-row packer tiles whole


def layer_shard(n_layers: int, world: int, rank: int) -> list[int]:
    """Layers owned by `rank` (round robin keeps the per-rank work within one layer of equal)."""
    return [l for l in range(n_layers) if l % world == rank]


def row_range(OC: int, world: int, rank: int) -> tuple[int, int]:
    if OC % (world * ROW_ALIGN):
        raise ValueError(f"out_features={OC} must be a multiple of world*{ROW_ALIGN}={world * ROW_ALIGN}")
    n = OC // world
    return rank * n, (rank + 1) * n


def shard_packed_rows(p: dict, world: int, rank: int) -> dict:
    """Slice a packed tensor set to the output rows of `rank` (contiguous copies)."""
    OC = p["weight"].shape[0]
    r0, r1 = row_range(OC, world, rank)
    return dict(weight=p["weight"][r0:r1].contiguous(), weight_last=p["weight_last"][r0:r1].contiguous(),
                zeros_and_scales=p["zeros_and_scales"][r0:r1].contiguous(),
                zeros_2nd=p["zeros_2nd"][r0 // 4:r1 // 4].contiguous(),
                scales_2nd=p["scales_2nd"][r0 // 4:r1 // 4].contiguous(),
                scales_4b=p["scales_4b"][r0:r1].contiguous(), zeros_4b=p["zeros_4b"][r0 // 8:r1 // 8].contiguous())


def gather_columns(y_local: torch.Tensor, group=None) -> torch.Tensor:
    """[M, N/W] per rank -> [M, N] on every rank (all-gather along the column dimension)."""
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    M, n = y_local.shape
    buf = torch.empty((world, M, n), dtype=y_local.dtype, device=y_local.device)
    try:
        dist.all_gather_into_tensor(buf, y_local.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):      # backends without the fused form
        dist.all_gather(list(buf.unbind(0)), y_local.contiguous(), group=group)
    return buf.permute(1, 0, 2).reshape(M, world * n)


class ColumnShardedMXQLinear:
    """y = x @ dequant(W)^T with W's output rows sharded over the process group."""

    def __init__(self, packed_local: dict, OC_total: int, group=None, mode: str = "nccl"):
        from . import ops
        self.ops = ops
        self.p = packed_local
        self.OC_total = OC_total
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.OC_local, self.IC = ops._packed_dims(packed_local)
        self.mode = mode
        self._ws = None
        self._ws_M = -1
        self._symm = None
        self._fused = None

    # -- fused path: the GEMM epilogue writes its tile into every peer's output over NVLink ------
    def _symm_out(self, M: int, device):
        """TWO symmetric [M, OC_total] buffers used by alternate calls.  Call k+1 stores into the
        other buffer, so a fast rank cannot overwrite a result a slower peer is still reading; by
        the time call k+2 reuses call k's buffer every rank has passed the barrier that ends call
        k+1, which on each rank's stream comes after its reads of call k's result."""
        if self._symm is not None and self._symm[0][0].shape[0] == M:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        bufs = []
        for _ in range(2):
            t = symm_mem.empty((M, self.OC_total), dtype=torch.float16, device=device)
            try:
                hdl = symm_mem.rendezvous(t, grp)
            except TypeError:
                hdl = symm_mem.rendezvous(t, grp.group_name)
            bufs.append((t, hdl))
        self._symm = bufs
        return self._symm

    def _fused_state(self, M: int, device):
        if self._fused is not None and self._fused["M"] == M:
            return self._fused
        import ctypes as C
        bufs = self._symm_out(M, device)
        self._fused = dict(M=M, pstruct=self.ops.L.packed_struct(self.p), parity=0, phased=None,
                           out=[t for t, _ in bufs], hdl=[h for _, h in bufs],
                           ptrs=[(C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs]) for _, h in bufs],
                           mc=[int(getattr(h, "multicast_ptr", 0) or 0) for _, h in bufs])
        return self._fused

    # -- phased exchange ---------------------------------------------------------------------------
    def _phase_plan(self, M: int):
        """(groups of (row0, rows), split) or None when the shard cannot be phased (rows not a multiple
        of the 256-row tile, or K not a multiple of 256)."""
        import os
        BN, BM2, PAIRS = 256, 512, 74
        if self.OC_local % BN or self.IC % 256 or self.OC_local < 2 * BN:
            return None
        nt, mt = self.OC_local // BN, -(-M // BM2)
        G = int(os.environ.get("MXQ_DIST_PHASES", "0")) or (2 if nt < 8 else 4)
        G = max(2, min(G, nt))
        base, extra = divmod(nt, G)
        groups, t0 = [], 0
        for g in range(G):
            n = base + (1 if g < extra else 0)
            groups.append((t0 * BN, n * BN))
            t0 += n
        tiles = max(n // BN for _, n in groups) * mt
        split = max(1, min(8, PAIRS // tiles, self.IC // 1024))     # a slice keeps >= 16 K blocks
        return groups, split

    def _phased_state(self, M: int, device):
        st = self._fused_state(M, device)
        if st.get("phased") is None:
            plan = self._phase_plan(M)
            if plan is None:
                st["phased"] = False
            else:
                import ctypes as C
                groups, split = plan
                L = self.ops.L
                need = max(L.lib().mxq_gemm_partials_workspace_bytes(M, rows, split) for _, rows in groups)
                sliced = []
                for row0, rows in groups:
                    q = dict(weight=self.p["weight"][row0:row0 + rows], weight_last=self.p["weight_last"][row0:row0 + rows],
                             zeros_and_scales=self.p["zeros_and_scales"][row0:row0 + rows],
                             zeros_2nd=self.p["zeros_2nd"][row0 // 4:(row0 + rows) // 4],
                             scales_2nd=self.p["scales_2nd"][row0 // 4:(row0 + rows) // 4],
                             scales_4b=self.p["scales_4b"][row0:row0 + rows], zeros_4b=self.p["zeros_4b"][row0 // 8:(row0 + rows) // 8])
                    sliced.append((L.packed_struct(q), row0, rows, q))   # row slices are contiguous views: no copies
                st["phased"] = dict(split=split, groups=sliced,
                                    ws=[torch.empty(int(need), dtype=torch.uint8, device=device) for _ in range(2)],
                                    side=torch.cuda.Stream(device=device),
                                    ev_part=[torch.cuda.Event() for _ in sliced], ev_red=[torch.cuda.Event() for _ in sliced],
                                    null_peers=(C.c_void_p * 1)(None))
        return st

    def _forward_phased(self, x: torch.Tensor, multicast: bool) -> torch.Tensor:
        L = self.ops.L
        M = x.shape[0]
        st = self._phased_state(M, x.device)
        ph = st["phased"]
        if not ph:
            return None
        b = st["parity"]
        st["parity"] = b ^ 1
        if multicast and not st["mc"][b]:
            raise RuntimeError("symmetric memory has no multicast mapping on this system (mode='mc2')")
        main = torch.cuda.current_stream(x.device)
        side = ph["side"]
        col_base = self.rank * self.OC_local
        lib = L.lib()
        with L.on(x):
            for g, (pstruct, row0, rows, _keep) in enumerate(ph["groups"]):
                ws = ph["ws"][g & 1]
                if g >= 2:
                    main.wait_event(ph["ev_red"][g - 2])          # the workspace of group g-2 has been reduced
                L.check(lib.mxq_gemm_partials(x.data_ptr(), pstruct, M, self.IC, rows, ph["split"], ws.data_ptr(), ws.numel(),
                                              main.cuda_stream), "mxq_gemm_partials")
                ph["ev_part"][g].record(main)
                side.wait_event(ph["ev_part"][g])
                L.check(lib.mxq_gemm_reduce_store(ws.data_ptr(), ws.numel(), None if multicast else st["ptrs"][b], 0 if multicast else self.world,
                                                  st["mc"][b] if multicast else None, M, rows, ph["split"], self.OC_total,
                                                  col_base + row0, side.cuda_stream), "mxq_gemm_reduce_store")
                ph["ev_red"][g].record(side)
            main.wait_stream(side)
        st["hdl"][b].barrier(channel=0)
        return st["out"][b]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ops = self.ops
        M = x.shape[0]
        if self._ws is None or self._ws_M != M:
            self._ws = ops.gemm_workspace(M, x.shape[1], self.OC_local, x.device)
            self._ws_M = M
        if self.world == 1:
            return ops.gemm(x, self.p, workspace=self._ws, validate=False)
        if self.mode in ("mc2", "p2p2") and self.world > 1:
            if not x.is_cuda or x.dtype != torch.float16 or not x.is_contiguous() or x.dim() != 2 or x.shape[1] != self.IC:
                raise ValueError(f"x must be a contiguous CUDA fp16 [M, {self.IC}] tensor")
            y = self._forward_phased(x, self.mode == "mc2")
            if y is not None:
                return y
            # shards that cannot be phased (fewer than two 256-row tiles) take the fused epilogue
        if self.mode in ("mc", "p2p", "mc2", "p2p2"):
            # argument marshalling is cached per M: at 8 ranks a shard's GEMM is ~100 us and the
            # Python call path must not be the longer one
            st = self._fused_state(M, x.device)
            L = ops.L
            if not x.is_cuda or x.dtype != torch.float16 or not x.is_contiguous() or x.dim() != 2 or x.shape[1] != self.IC:
                raise ValueError(f"x must be a contiguous CUDA fp16 [M, {self.IC}] tensor")
            col0 = self.rank * self.OC_local
            b = st["parity"]
            st["parity"] = b ^ 1
            with L.on(x) as stream:
                if self.mode in ("mc", "mc2"):
                    if not st["mc"][b]:
                        raise RuntimeError("symmetric memory has no multicast mapping on this system (mode='mc')")
                    rc = L.lib().mxq_gemm_multicast(x.data_ptr(), st["pstruct"], st["mc"][b], M, self.IC, self.OC_local,
                                                    self.OC_total, col0, self._ws.data_ptr(), self._ws.numel(), stream)
                else:
                    rc = L.lib().mxq_gemm_scatter(x.data_ptr(), st["pstruct"], st["ptrs"][b], self.world, M, self.IC,
                                                  self.OC_local, self.OC_total, col0, self._ws.data_ptr(),
                                                  self._ws.numel(), stream)
            L.check(rc, "mxq_gemm_" + ("multicast" if self.mode in ("mc", "mc2") else "scatter"))
            st["hdl"][b].barrier(channel=0)    # every rank's tiles have landed in every rank's buffer
            return st["out"][b]
        y_local = ops.gemm(x, self.p, workspace=self._ws, validate=False)
        return gather_columns(y_local, self.group)

    __call__ = forward
