"""LLM-QAT/models/utils_quant.py names (:31-199, :310-475, :601-727) backed by the sm_100a kernels."""
from mxq_b200.utils_quant import AsymQuantizer, MXAsymQuantizer, QuantizeLinear, SymQuantizer  # noqa: F401
