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
    from torch.utils import cpp_extension as ce
    if os.path.exists(SO) and not force:
        if not os.path.exists(os.path.join(OUT, "awq_gemm_ref.so")):
            inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
            build_awq_gemm(ce, inc, [f"-L{p}" for p in ce.library_paths(device_type="cuda")])
        return SO
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
    build_awq_gemm(ce, inc, libs)
    return SO


def build_awq_gemm(ce, inc, libs):
    """Second module: the reference's AWQ 4-bit GEMM (gemm_cuda_gen.cu) behind oracle/awq_gemm_ref_binding.cpp."""
    so = os.path.join(OUT, "awq_gemm_ref.so")
    common = ["-O3", "-std=c++17", "-DTORCH_EXTENSION_NAME=awq_gemm_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1", "-DENABLE_BF16", f"-I{SRC}"]
    o1, o2 = os.path.join(OUT, "gemm_cuda_gen.o"), os.path.join(OUT, "awq_binding.o")
    a = subprocess.Popen(["nvcc", *common, *inc, "-gencode", "arch=compute_100a,code=sm_100a", "--use_fast_math",
                          "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "--expt-relaxed-constexpr",
                          "-Xcompiler", "-fPIC", "-c", os.path.join(SRC, "quantization", "gemm_cuda_gen.cu"), "-o", o1])
    b = subprocess.Popen(["g++", *common, *inc, "-fPIC", "-c", os.path.join(HERE, "awq_gemm_ref_binding.cpp"), "-o", o2])
    if a.wait() != 0 or b.wait() != 0:
        print("reference AWQ GEMM failed to compile (tests that need it skip)")
        return None
    subprocess.check_call(["g++", "-shared", "-o", so, o1, o2, *libs, "-lc10", "-ltorch", "-ltorch_cpu",
                           "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
                           "-Wl,-rpath," + ce.library_paths(device_type="cuda")[0]])
    os.remove(o1)
    os.remove(o2)
    print(so)
    return so


if __name__ == "__main__":
    main(force="--force" in sys.argv)
