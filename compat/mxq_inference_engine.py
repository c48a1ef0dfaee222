"""Drop-in module name of the reference's pybind extension (cuda_kernel/csrc/pybind.cpp:6-10)."""
from mxq_b200.engine import (gemm_forward_cuda, gemm_mxq_forward_cuda, gemv_forward_cuda,  # noqa: F401
                             gemv_mxq_forward_cuda)
