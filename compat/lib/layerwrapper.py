from mxq_b200.layerwrapper import WrappedGPT  # noqa: F401  (mxq_quant/lib/layerwrapper.py:5-35)
