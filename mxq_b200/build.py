"""Build libmxq_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mxq_b200.build [--force] [--verbose]

The quantizer kernels are compiled WITHOUT --use_fast_math and with -fmad=false: bit-exact
parity with the reference needs IEEE division and no FMA contraction.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmxq_b200.so")

# file -> extra flags
SOURCES = {
    "misc.cu": [],
    "fakequant.cu": ["-fmad=false"],
    "actquant.cu": ["-fmad=false"],
    "calib.cu": [],
    "ptq.cu": ["-fmad=false"],
    "pack.cu": ["-fmad=false"],
    "gemv.cu": [],
    "gemv_mma.cu": [],
    "gemv_chain.cu": [],
    "gemm_tcgen05.cu": [],
    "awq_gemm.cu": ["-fmad=false"],
}

COMMON = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps(src: str):
    return [src, os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "mxq_b200.h"),
            os.path.abspath(__file__)]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    common = list(COMMON)
    if os.environ.get("MXQ_DEBUG") == "1":      # watchdog traps + pinned-host records in the GEMM barriers
        common.append("-DMXQ_DEBUG")
    if os.environ.get("MXQ_CHAIN_TRACE") == "1":  # clock64 stamps in the persistent GEMV chain (profiles/r2_gemv_persistent_trace.py)
        common.append("-DMXQ_CHAIN_TRACE")
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    objs = []
    for name, extra in SOURCES.items():
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, _deps(src)):
            cmd = [nvcc, *common, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stdout + r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
