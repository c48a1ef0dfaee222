"""mxq_quant/lib/prune.py names imported by mxq_quant/main.py:8.  Only the mxq entry is on the hot
path; the pruning variants (unreachable with the default --sparsity_ratio 0, main.py:64-74) raise."""
import torch
import torch.nn as nn

from mxq_b200.prune import find_layers, nas_quant  # noqa: F401


def check_sparsity(model):
    """prune.py:38-62: fraction of exactly-zero weights per decoder layer and overall."""
    use_cache = model.config.use_cache
    model.config.use_cache = False
    count, total = 0, 0
    for i, layer in enumerate(model.model.layers):
        sub_count, sub_params = 0, 0
        for name, m in find_layers(layer).items():
            W = m.weight.data
            sub_count += (W == 0).sum().item()
            sub_params += W.numel()
        count += sub_count
        total += sub_params
        print(f"layer {i} sparsity {float(sub_count) / max(sub_params, 1):.6f}")
    model.config.use_cache = use_cache
    return float(count) / max(total, 1)


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is a pruning method outside the MXQ hot path (use the reference's lib.prune)")
    return f


prune_wanda = _not_on_path("prune_wanda")
prune_magnitude = _not_on_path("prune_magnitude")
prune_sparsegpt = _not_on_path("prune_sparsegpt")
prune_ablate = _not_on_path("prune_ablate")
