"""Compile the REFERENCE CUDA extension (mxq_inference_engine: gemv_forward_cuda,
gemv_mxq_forward_cuda) from its own sources where they lie under /root/reference into
oracle/_ref/ -- TEST INFRASTRUCTURE: a same-box GPU comparator and a second pin for the packed
layout's bit order.  Nothing is copied; outputs only go to oracle/_ref/ (git-ignored, shipped to
the GPU box with the snapshot).  Not the reference's own build system: three nvcc/g++ commands.

    python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MXQ_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "mxq_quant", "cuda_kernel", "csrc")
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "mxq_inference_engine.so")


def main(force: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        print("reference sources not present; skipping oracle/_ref build")
        return None
    if os.path.exists(SO) and not force:
        return SO
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = ["-O3", "-std=c++17", "-DTORCH_EXTENSION_NAME=mxq_inference_engine", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1", "-DENABLE_BF16"]
    nvcc = ["nvcc", *common, *inc, "-gencode", "arch=compute_100a,code=sm_100a", "--use_fast_math",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_BFLOAT16_OPERATORS__",
            "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "--expt-relaxed-constexpr", "--expt-extended-lambda",
            "-Xcompiler", "-fPIC", "-c"]
    objs = []
    jobs = []
    for rel in ("quantization/gemv_cuda.cu", "quantization/gemv_mxq_cuda.cu"):
        o = os.path.join(OUT, os.path.basename(rel).replace(".cu", ".o"))
        objs.append(o)
        jobs.append(subprocess.Popen([*nvcc, os.path.join(SRC, rel), "-o", o]))
    o = os.path.join(OUT, "pybind.o")
    objs.append(o)
    jobs.append(subprocess.Popen(["g++", *common, *inc, "-fPIC", "-c", os.path.join(SRC, "pybind.cpp"), "-o", o]))
    for j in jobs:
        if j.wait() != 0:
            raise SystemExit("reference extension failed to compile")
    libs = [f"-L{p}" for p in ce.library_paths(device_type="cuda")]
    subprocess.check_call(["g++", "-shared", "-o", SO, *objs, *libs, "-lc10", "-ltorch", "-ltorch_cpu",
                           "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
                           "-Wl,-rpath," + ce.library_paths(device_type="cuda")[0]])
    for o in objs:
        os.remove(o)
    print(SO)
    return SO


if __name__ == "__main__":
    main(force="--force" in sys.argv)
