"""ctypes loader for libmxq_b200.so (the C ABI declared in include/mxq_b200.h).

There is no CPU fallback: if the library is missing or a call is made with non-CUDA tensors the
wrappers raise.  Build with ``python -m mxq_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmxq_b200.so")

MXQ_F32, MXQ_F16, MXQ_BF16 = 0, 1, 2
POOL = 0x80

_DTYPE = {torch.float32: MXQ_F32, torch.float16: MXQ_F16, torch.bfloat16: MXQ_BF16}

# every symbol include/mxq_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "mxq_version", "mxq_error_string", "mxq_fakequant_fwd", "mxq_fakequant_fwd_multi", "mxq_ste_bwd",
    "mxq_segquant_workspace_bytes", "mxq_segquant_fwd",
    "mxq_colsumsq_workspace_bytes", "mxq_colsumsq", "mxq_colsumsq_ex", "mxq_wanda_metric",
    "mxq_allocate_bits_workspace_bytes", "mxq_allocate_bits",
    "mxq_ptq_workspace_bytes", "mxq_ptq_quant", "mxq_rowquant",
    "mxq_pack_workspace_bytes", "mxq_pack", "mxq_ptq_quant_pack", "mxq_unpack", "mxq_gemv", "mxq_gemv_ex", "mxq_gemv_grouped", "mxq_gemv_grouped_perm", "mxq_gemv_chain_plan_bytes", "mxq_gemv_chain_plan", "mxq_gemv_chain_run", "mxq_gather_groups", "mxq_awq_gemv", "mxq_awq_gemm_workspace_bytes", "mxq_awq_gemm",
    "mxq_gemm_workspace_bytes", "mxq_gemm", "mxq_gemm_plan", "mxq_gemm_scatter", "mxq_gemm_multicast", "mxq_gemm_dense",
    "mxq_gemm_partials_workspace_bytes", "mxq_gemm_partials", "mxq_gemm_reduce_store",
]


class PackedC(C.Structure):
    """mxq_packed_t"""
    _fields_ = [("weight", C.c_void_p), ("weight_last", C.c_void_p),
                ("zeros_and_scales", C.c_void_p), ("zeros_2nd", C.c_void_p),
                ("scales_2nd", C.c_void_p), ("scales_4b", C.c_void_p), ("zeros_4b", C.c_void_p)]


class GemvJobC(C.Structure):
    """mxq_gemv_job_t"""
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("w", PackedC), ("IC", C.c_int64), ("OC", C.c_int64),
                ("dep", C.c_int32), ("reserved", C.c_int32)]


GEMV_CHAIN_MAX_JOBS = 64
GEMV_CHAIN_SYNC_WORDS = GEMV_CHAIN_MAX_JOBS + 1
GEMV_CHAIN_PDL = 1                 # mxq_gemv_chain_run flag (MXQ_GEMV_CHAIN_PDL)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the mxq_b200 CUDA library is not built and there is no CPU "
            "fallback. Run `python -m mxq_b200.build`.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t
    L.mxq_version.restype = i32
    L.mxq_error_string.restype = C.c_char_p
    L.mxq_error_string.argtypes = [i32]
    L.mxq_fakequant_fwd.argtypes = [vp, vp, vp, i64, i64, i32, i32, i32, vp, vp]
    L.mxq_fakequant_fwd_multi.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, i64, i32, i32, i32, vp, vp]
    L.mxq_ste_bwd.argtypes = [vp, vp, vp, i64, i32, f32, f32, vp]
    L.mxq_segquant_workspace_bytes.restype = sz
    L.mxq_segquant_workspace_bytes.argtypes = [i64, i64, i32]
    L.mxq_segquant_fwd.argtypes = [vp, vp, i64, i64, i32, i32, i32, i64, i64, vp, sz, vp]
    L.mxq_colsumsq_workspace_bytes.restype = sz
    L.mxq_colsumsq_workspace_bytes.argtypes = [i64, i64]
    L.mxq_colsumsq.argtypes = [vp, i64, i64, i32, vp, f32, f32, i32, vp, sz, vp]
    L.mxq_colsumsq_ex.argtypes = [vp, i64, i64, i32, vp, f32, f32, i32, i32, vp, sz, vp]
    L.mxq_wanda_metric.argtypes = [vp, vp, vp, i64, i64, i32, vp]
    L.mxq_allocate_bits_workspace_bytes.restype = sz
    L.mxq_allocate_bits_workspace_bytes.argtypes = [i64]
    L.mxq_allocate_bits.argtypes = [vp, vp, i64, i64, i32, i32, vp, vp, vp, sz, vp]
    L.mxq_ptq_workspace_bytes.restype = sz
    L.mxq_ptq_workspace_bytes.argtypes = [i64, i64]
    L.mxq_ptq_quant.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, sz, vp]
    L.mxq_rowquant.argtypes = [vp, vp, vp, vp, vp, i64, i64, i32, i32, vp]
    L.mxq_pack_workspace_bytes.restype = sz
    L.mxq_pack_workspace_bytes.argtypes = [i64, i64]
    L.mxq_pack.argtypes = [vp, vp, i64, i64, PackedC, vp, sz, vp]
    L.mxq_ptq_quant_pack.argtypes = [vp, vp, vp, vp, i64, i64, PackedC, vp, sz, vp]
    L.mxq_unpack.argtypes = [PackedC, i64, i64, vp, i32, vp]
    L.mxq_gemv.argtypes = [vp, PackedC, vp, i64, i64, i64, vp]
    L.mxq_gemv_ex.argtypes = [vp, PackedC, vp, i64, i64, i64, C.c_uint, vp]
    L.mxq_gemv_grouped.argtypes = [vp, C.POINTER(PackedC), C.POINTER(vp), i32, i64, i64, i64, C.c_uint, vp]
    L.mxq_gemv_grouped_perm.argtypes = [vp, C.POINTER(PackedC), C.POINTER(vp), i32, i64, i64, i64, vp, C.c_uint, vp]
    L.mxq_gemv_chain_plan_bytes.restype = sz
    L.mxq_gemv_chain_plan_bytes.argtypes = []
    L.mxq_gemv_chain_plan.argtypes = [C.POINTER(GemvJobC), i32, vp]
    L.mxq_gemv_chain_run.argtypes = [vp, vp, vp, C.c_uint, vp]
    L.mxq_gather_groups.argtypes = [vp, vp, vp, i64, i64, vp]
    L.mxq_awq_gemv.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i32, vp]
    L.mxq_awq_gemm_workspace_bytes.restype = sz
    L.mxq_awq_gemm_workspace_bytes.argtypes = [i64, i64]
    L.mxq_awq_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i32, vp, sz, vp]
    L.mxq_gemm_workspace_bytes.restype = sz
    L.mxq_gemm_workspace_bytes.argtypes = [i64, i64, i64]
    L.mxq_gemm.argtypes = [vp, PackedC, vp, i64, i64, i64, vp, sz, vp]
    L.mxq_gemm_dense.argtypes = [vp, vp, vp, i64, i64, i64, vp]
    L.mxq_gemm_scatter.argtypes = [vp, PackedC, C.POINTER(vp), i32, i64, i64, i64, i64, i64, vp, sz, vp]
    L.mxq_gemm_plan.argtypes = [i64, i64, i64, i32, C.POINTER(C.c_int32), C.POINTER(sz)]
    L.mxq_gemm_partials_workspace_bytes.restype = sz
    L.mxq_gemm_partials_workspace_bytes.argtypes = [i64, i64, i32]
    L.mxq_gemm_partials.argtypes = [vp, PackedC, i64, i64, i64, i32, vp, sz, vp]
    L.mxq_gemm_reduce_store.argtypes = [vp, sz, C.POINTER(vp), i32, vp, i64, i64, i32, i64, i64, vp]
    L.mxq_gemm_multicast.argtypes = [vp, PackedC, vp, i64, i64, i64, i64, i64, vp, sz, vp]
    _lib = L
    return L


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().mxq_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def dtype_enum(t: torch.Tensor) -> int:
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"mxq_b200: unsupported dtype {t.dtype}") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    """Every tensor of a call must live on ONE CUDA device (the kernels take raw pointers)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mxq_b200 kernels need CUDA tensors (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"mxq_b200: tensors of one call live on different devices ({dev} and {t.device})")


class on:
    """`with on(tensor) as stream:` -- makes the tensor's device current for the launch (the reference
    model may be spread over GPUs by `hf_device_map`, mxq_quant/lib/prune.py:371-378) and yields
    torch's current stream ON THAT DEVICE; a no-op switch when it already is the current device."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("mxq_b200 kernels need CUDA tensors (no CPU fallback)")
        self.idx = t.device.index

    def __enter__(self) -> int:
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return torch.cuda.current_stream(self.idx).cuda_stream

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def packed_struct(p: dict) -> PackedC:
    return PackedC(p["weight"].data_ptr(), p["weight_last"].data_ptr(),
                   p["zeros_and_scales"].data_ptr(), p["zeros_2nd"].data_ptr(),
                   p["scales_2nd"].data_ptr(), p["scales_4b"].data_ptr(), p["zeros_4b"].data_ptr())
