"""CPU: bench.py's measurement contract, checked on CODE (not on committed result files): the
reference arm runs end to end on a reduced sample and prints the full JSON line; the helper that turns
a committed ncu capture into `roofline.traffic` parses the real CSV; the default run refuses to start
without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=dict(os.environ, **(env or {})), timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"],
             {"MXQ_BENCH_REF_TOKENS": "256", "MXQ_BENCH_REF_ROWS": "0.02"})
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mxq_quant_pass_hbm_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["sample"] and abs(c["value"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ncu_traffic_lookup_reads_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic("colsumsq_partial_kernel", "stats_4096")
    assert t is not None and "r2_ncu_traffic.csv" in src
    # the capture of the SHIPPED instantiation: <__half, 16, 2>, 128 x 2048 tokens x 4096 channels x 2 B read once
    assert "16, 2" in src and 0.999 < t / (128 * 2048 * 4096 * 2) < 1.01
    # the persistent decode chain: 32 x 4096^2 jobs of 6,318,080 packed + activation bytes each, read once
    t, src = bench.ncu_traffic("gemv_chain_kernel", "gemv_chain_32x4096x4096")
    assert t is not None and 0.99 < t / (32 * 6318080) < 1.06
    # bench.py's own chain: the 56 linears of 8 layers (packed tensors + activations + outputs), read once
    from mxq_b200.prune import packed_nbytes
    algo = 8 * sum(packed_nbytes(oc, ic) + 2 * (oc + ic) for oc, ic in [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)])
    t, src = bench.ncu_traffic("gemv_chain_kernel", "gemv_chain_56linear")
    assert t is not None and 0.99 < t / algo < 1.06
    assert bench.ncu_traffic("no_such_kernel")[0] is None
    pk = bench.peaks()
    assert pk["hbm"] > 1000 and pk["tf_burst"] > 100 and pk["src"] in ("measured", "fallback")
    cfg = bench.workload_config(4)
    assert "layer % 4" in cfg["sharding"] and cfg["layers"] == 32 and cfg["nsamples"] == 128


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "3", "--no-components", "--no-e2e"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
