"""Drop-in for the pybind module ``mxq_inference_engine`` (mxq_quant/cuda_kernel/csrc/pybind.cpp:6-10):
``gemv_mxq_forward_cuda`` (gemv_mxq_cuda.h:4-12) and ``gemv_forward_cuda`` (gemv_cuda.h:4-9), plus
the prefill ``gemm_mxq_forward_cuda`` the reference lacks.  Differences from the reference
binding: any in_features % 64 == 0 (not only 4096), shape/dtype/device are checked, and the
kernels run on torch's current stream instead of the legacy default stream."""
from __future__ import annotations


from . import ops


def _as_packed(kernel, kernel_last, zeros_and_scales, scales_2nd, zeros_2nd, scales_4b, zeros_4b):
    return dict(weight=kernel, weight_last=kernel_last, zeros_and_scales=zeros_and_scales,
                zeros_2nd=zeros_2nd, scales_2nd=scales_2nd, scales_4b=scales_4b, zeros_4b=zeros_4b)


def gemv_mxq_forward_cuda(in_feats, kernel, kernel_last, zeros_and_scales, scales_2nd, zeros_2nd,
                          scales_4b, zeros_4b, group_size):
    """in_feats fp16 [B, IC] -> fp16 [B, OC] (gemv_mxq_cuda.cu:225-273)."""
    if group_size != 16:
        # the reference launches nothing and returns uninitialised memory (gemv_mxq_cuda.cu:263-272)
        raise ValueError("gemv_mxq_forward_cuda: only group_size == 16 exists")
    p = _as_packed(kernel, kernel_last, zeros_and_scales, scales_2nd, zeros_2nd, scales_4b, zeros_4b)
    return ops.gemv(in_feats, p)


def gemm_mxq_forward_cuda(in_feats, kernel, kernel_last, zeros_and_scales, scales_2nd, zeros_2nd,
                          scales_4b, zeros_4b, group_size=16):
    """Prefill counterpart on tcgen05/TMEM: in_feats fp16 [M, IC] -> fp16 [M, OC]."""
    if group_size != 16:
        raise ValueError("gemm_mxq_forward_cuda: only group_size == 16 exists")
    p = _as_packed(kernel, kernel_last, zeros_and_scales, scales_2nd, zeros_2nd, scales_4b, zeros_4b)
    return ops.gemm(in_feats, p)


def gemv_forward_cuda(in_feats, kernel, scaling_factors, zeros, group_size):
    """AWQ uniform 4-bit GEMV (gemv_cuda.cu:346-399)."""
    return ops.awq_gemv(in_feats, kernel, scaling_factors, zeros, group_size)


def gemm_forward_cuda(in_feats, kernel, scaling_factors, zeros, split_k_iters=1):
    """AWQ uniform 4-bit prefill GEMM (gemm_cuda_gen.cu:424-478, gemm_cuda.h; present in the reference's
    sources but absent from its build, setup.py:37-41).  `split_k_iters` is accepted and ignored: the
    tcgen05 kernel accumulates all of K in fp32 (the reference adds fp16 split-K partials, :477)."""
    return ops.awq_gemm(in_feats, kernel, scaling_factors, zeros)
