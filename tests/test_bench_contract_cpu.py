"""CPU: the JSON lines bench.py printed on the B200 boxes (committed under profiles/) carry every key
of the measurement contract, with consistent values -- a guard against a bench edit dropping one."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches"]


def _line(path):
    with open(path) as f:
        rows = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(rows[-1])


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        pytest.skip("no committed bench line " + pattern)
    return files[-1]


def test_default_workload_line():
    d = _line(_latest("r1_bench_n1_v*.json"))
    for k in BASE + ["cpu_baseline", "clocks", "components"]:
        assert k in d, k
    assert d["metric"] == "mxq_quant_pass_hbm_GBps" and d["unit"] == "GB/s" and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= 0.99 * r["bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # host buffers can only be slower
    assert d["gpu_launches"] > 0
    # value = whole-job bytes / device time
    assert abs(d["value"] - d["job_bytes_per_step"] / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line(_latest("r1_bench_reference_arm*.json"))
    assert d["impl"] == "reference" and d["metric"] == "mxq_quant_pass_hbm_GBps"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


@pytest.mark.parametrize("pattern,metric", [("r1_bench_gemm70b_n8_v*.json", "mxq_dequant_gemm_70b_TFLOPs"),
                                            ("r1_bench_gemm70b_n2_v*.json", "mxq_dequant_gemm_70b_TFLOPs"),
                                            ("r1_bench_qat_n2_v*.json", "qat_step_tokens_per_s"),
                                            ("r1_bench_ptq_n2_v*.json", "mxq_quant_pass_hbm_GBps")])
def test_multi_gpu_lines(pattern, metric):
    d = _line(_latest(pattern))
    assert d["metric"] == metric and d["n_gpus"] >= 2 and d["value"] > 0
    assert d["scaling"] in ("weak", "strong")
    if metric.startswith("mxq_dequant"):
        # the reported exchange is the fastest mode that ran on every shape
        modes = {m: d[m + "_TFLOPs"] for m in ("nccl", "p2p", "mc") if d.get(m + "_TFLOPs")}
        assert d["config"]["exchange"] == max(modes, key=modes.get)
        assert abs(d["value"] - modes[d["config"]["exchange"]]) < 1e-6 * d["value"]
        assert d["value"] <= d["gemm_only_TFLOPs"]
