"""mxq_b200: B200-native (sm_100a) implementation of the MXQ quantization hot path.

Host-side mirrors of the reference interfaces (same names / arguments):
  utils_quant.MXAsymQuantizer, QuantizeLinear, SymQuantizer, AsymQuantizer (LLM-QAT/models/utils_quant.py)
  mxqgpt.MXQGPT, prune.nas_quant, layerwrapper.WrappedGPT, quantizer.Quantizer (mxq_quant/lib)
  engine.gemv_mxq_forward_cuda / gemv_forward_cuda          (mxq_quant/cuda_kernel)
  packed_linear.MXQLinear / save_packed / load_packed       (packed-checkpoint consumer, no reference analogue)
All computation happens in libmxq_b200.so through the C ABI in include/mxq_b200.h.
"""
from . import ops  # noqa: F401
from .utils_quant import MXAsymQuantizer, QuantizeLinear, SymQuantizer, AsymQuantizer  # noqa: F401
from .mxqgpt import MXQGPT  # noqa: F401
from .layerwrapper import WrappedGPT  # noqa: F401
from .quantizer import Quantizer  # noqa: F401
from .packed_linear import MXQLinear  # noqa: F401

__version__ = "0.1.0"
