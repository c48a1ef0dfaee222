"""Debugging aid: list the golden actquant cases whose GPU output differs, with the first mismatches."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from mxq_b200 import SymQuantizer, AsymQuantizer
from tests.gpu_util import to_dev, to_np
aq = np.load("tests/golden/actquant.npz")
clip = torch.tensor([-2.0, 2.0]); dev = torch.device("cuda:0")
for k in sorted(k for k in aq.files if k.endswith("/y")):
    mode, dtype, case, b, _ = k.split("/")
    fn = SymQuantizer if mode == "sym" else AsymQuantizer
    xs = aq[f"{mode}/{dtype}/{case}/x"]
    y = to_np(fn.apply(to_dev(xs, dtype, dev), clip, int(b[1:]), case.startswith("layerwise")))
    ref = aq[k]
    bad = (y.view(np.uint32) != ref.view(np.uint32)) & ~(np.isnan(y) & np.isnan(ref))
    if bad.any():
        idx = np.argwhere(bad)[:3]
        print(k, int(bad.sum()), "of", bad.size, [(tuple(int(j) for j in i), float(xs[tuple(i)]), float(y[tuple(i)]), float(ref[tuple(i)])) for i in idx])
print("done")
