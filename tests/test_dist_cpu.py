"""CPU (gloo, world_size 2): the sharding logic of the multi-GPU paths -- layer round-robin for the
PTQ pass and row-sharded packed tensors + column all-gather for the 70B-shape GEMM.  The per-rank
compute is done by the oracle here (no GPU); the GPU kernels run the same shards on the box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mxq_b200 import dist as mdist
from oracle import mxq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) PTQ pass: every layer quantized exactly once, results identical to a serial run
        n_layers, shape = 6, (32, 128)
        mine = mdist.layer_shard(n_layers, world, rank)
        sums = torch.zeros(n_layers, dtype=torch.float64)
        for l in mine:
            rng = np.random.default_rng(l)
            W = (rng.standard_normal(shape) * 0.02).astype(np.float16)
            sums[l] = float(O.fasterquant(W).astype(np.float64).sum())
        dist.all_reduce(sums)
        # (2) column-sharded GEMM + all-gather
        OC, IC, M = 64, 128, 5
        p = O.random_packed(OC, IC, seed=3)
        pt = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}
        local = mdist.shard_packed_rows(pt, world, rank)
        x = np.random.default_rng(9).standard_normal((M, IC)).astype(np.float16)
        y_local = O.gemm_mxq_f32(x, {k: v.numpy() for k, v in local.items()})
        y = mdist.gather_columns(torch.from_numpy(y_local), None)
        if rank == 0:
            q.put((sums.tolist(), y.numpy(), O.gemm_mxq_f32(x, p)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    sums, y, y_ref = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = []
    for l in range(6):
        W = (np.random.default_rng(l).standard_normal((32, 128)) * 0.02).astype(np.float16)
        want.append(float(O.fasterquant(W).astype(np.float64).sum()))
    assert sums == want
    assert np.array_equal(y, y_ref)


def test_shard_helpers():
    assert mdist.layer_shard(32, 8, 3) == [3, 11, 19, 27]
    cover = sorted(sum((mdist.layer_shard(32, 8, r) for r in range(8)), []))
    assert cover == list(range(32))
    assert mdist.row_range(8192, 8, 7) == (7168, 8192)
    with pytest.raises(ValueError):
        mdist.row_range(8200, 8, 0)
    p = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in O.random_packed(64, 128, seed=1).items()}
    parts = [mdist.shard_packed_rows(p, 2, r) for r in range(2)]
    full = O.decode_mxq({k: v.numpy() for k, v in p.items()})
    got = np.concatenate([O.decode_mxq({k: v.numpy() for k, v in s.items()}) for s in parts], axis=0)
    assert np.array_equal(full, got)      # per-row, 4-row and 8-row metadata travel with their rows
