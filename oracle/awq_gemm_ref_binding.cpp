// TEST INFRASTRUCTURE: exports the reference's AWQ 4-bit GEMM (gemm_forward_cuda, declared in
// mxq_quant/cuda_kernel/csrc/quantization/gemm_cuda.h, defined in gemm_cuda_gen.cu:424-478), which
// the reference's own pybind.cpp / setup.py leave out, so tests can pin mxq_awq_gemm's bit order and
// arithmetic against the unmodified kernel.  Built by oracle/build_ref.py into oracle/_ref/.
#include <torch/extension.h>

#include "quantization/gemm_cuda.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) { m.def("gemm_forward_cuda", &gemm_forward_cuda, "AWQ 4-bit GEMM (reference)"); }
