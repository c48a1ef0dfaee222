from mxq_b200.mxqgpt import MXQGPT  # noqa: F401  (mxq_quant/lib/mxqgpt.py:353-452)
