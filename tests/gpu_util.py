import numpy as np
import torch

TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def to_dev(a: np.ndarray, dtype: str, dev):
    """fp32 numpy array holding dtype-representable values -> device tensor of that dtype."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(TD[dtype])


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-for-bit equality; NaNs must sit in the same places (payload/sign of a NaN is not part
    of the contract: 0/0 in a constant fp16 group is NaN in the reference too)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def packed_to_dev(p: dict, dev):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in p.items()}


def packed_to_np(p: dict):
    return {k: v.detach().cpu().numpy() for k, v in p.items()}
