"""`lib` as mxq_quant names it.  prune / mxqgpt / quantizer / layerwrapper come from mxq_b200; data,
eval and the pruning variants are looked up in $MXQ_REFERENCE_ROOT/mxq_quant/lib when that is set."""
import os

_ref = os.environ.get("MXQ_REFERENCE_ROOT")
if _ref:
    _p = os.path.join(_ref, "mxq_quant", "lib")
    if os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
