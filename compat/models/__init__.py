"""`models` as LLM-QAT names it.  utils_quant comes from mxq_b200; everything else (modeling_llama_quant,
configuration_llama) is looked up in $MXQ_REFERENCE_ROOT/LLM-QAT/models when that is set."""
import os

_ref = os.environ.get("MXQ_REFERENCE_ROOT")
if _ref:
    _p = os.path.join(_ref, "LLM-QAT", "models")
    if os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
