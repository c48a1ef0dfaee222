from mxq_b200.quantizer import Quantizer  # noqa: F401  (mxq_quant/lib/quantizer.py:23-180)
