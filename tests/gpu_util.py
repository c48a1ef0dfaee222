import numpy as np
import torch

TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def to_dev(a: np.ndarray, dtype: str, dev):
    """fp32 numpy array holding dtype-representable values -> device tensor of that dtype."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(TD[dtype])


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def packed_to_dev(p: dict, dev):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in p.items()}


def packed_to_np(p: dict):
    return {k: v.detach().cpu().numpy() for k, v in p.items()}
