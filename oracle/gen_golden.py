"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python on seeded inputs.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box); the fixtures it writes are committed, together with this script, so that the oracle
and the CUDA path can be pinned to the reference everywhere else.

    python oracle/gen_golden.py            # rewrites tests/golden/

Reference entry points executed (nothing is copied from them):
  LLM-QAT/models/utils_quant.py:310-475   MXAsymQuantizer.forward / backward
  mxq_quant/lib/mxqgpt.py:353-452         MXQGPT.add_batch / fasterquant(blocksize=16)
  mxq_quant/lib/layerwrapper.py:22-35     WrappedGPT.add_batch
  mxq_quant/lib/quantizer.py:23-180       Quantizer (through fasterquant, and directly)
  LLM-QAT/models/utils_quant.py:31-199    SymQuantizer / AsymQuantizer forward / backward
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("MXQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, os.path.join(REF, "LLM-QAT", "models"))
    sys.path.insert(0, os.path.join(REF, "mxq_quant"))
    import utils_quant  # noqa
    from lib import mxqgpt, layerwrapper, quantizer  # noqa
    return utils_quant, mxqgpt, layerwrapper, quantizer


TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def fakequant_inputs():
    """name -> fp32 tensor [N, K] (rounded to the case dtype by the caller)."""
    cases = {}
    g = torch.Generator().manual_seed(0)
    cases["randn_64x256"] = torch.randn(64, 256, generator=g) * 0.02
    cases["randn_8x11008"] = torch.randn(8, 11008, generator=g) * 0.02
    x = torch.randn(16, 128, generator=g) * 0.02
    x[0, :16] = 0.125            # constant low-bit group: alpha == 0 -> divide by 1e-8
    x[1, 48:64] = -0.5           # 4-bit columns of row 1 all equal in one block, not in the other
    x[2, :] = 0.0                # whole row zero: every alpha == 0, pooled alpha == 0
    x[3, 0] = 2.5
    x[3, 1] = -2.0               # the STE clip plants of SURVEY 8d
    x[4, :] = torch.linspace(-1, 1, 128)
    x[5, :16] = torch.tensor([1e-9 * i for i in range(16)])   # alpha ~ 1.5e-8 comparable to 1e-8
    x[6, :] = x[6, :] * 1e-4     # alpha < 0.25 territory where +1e-8 moves fp32 alpha by ulps
    x[7, :] = x[7, :] * 50.0
    cases["edges_16x128"] = x
    cases["sin_2x128"] = (torch.sin(0.37 * torch.arange(256, dtype=torch.float32)).reshape(2, 128) * 0.05)
    cases["empty_0x64"] = torch.zeros(0, 64)
    return cases


def gen_fakequant(utils_quant):
    out = {}
    clip = torch.tensor([-2.0, 2.0])
    for dname, td in TD.items():
        for cname, x32 in fakequant_inputs().items():
            x = x32.to(td).clone().requires_grad_(True)
            y = utils_quant.MXAsymQuantizer.apply(x, clip, 2, False)
            g = torch.Generator().manual_seed(1)
            go = torch.randn(x.shape, generator=g).to(td)
            if x.numel():
                y.backward(go)
                gi = x.grad
            else:
                gi = torch.zeros_like(x)
            key = f"{dname}/{cname}"
            out[key + "/x"] = x.detach().float().numpy()
            out[key + "/y"] = y.detach().float().numpy()
            out[key + "/go"] = go.float().numpy()
            out[key + "/gi"] = gi.float().numpy()
    # w_bits = 3 and 4 low groups (QuantizeLinear passes w_bits as num_bits, utils_quant.py:637)
    for nb in (3, 4):
        x32 = fakequant_inputs()["randn_64x256"]
        y = utils_quant.MXAsymQuantizer.apply(x32.clone(), clip, nb, False)
        out[f"fp32/bits{nb}_64x256/x"] = x32.numpy()
        out[f"fp32/bits{nb}_64x256/y"] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "fakequant.npz"), **out)
    print("fakequant.npz", len(out), "arrays")


def gen_fasterquant(mxqgpt_mod, layerwrapper):
    torch.cuda.synchronize = lambda *a, **k: None      # mxqgpt.py:445 calls it unconditionally
    out = {}
    cases = {}
    g = torch.Generator().manual_seed(2)
    cases["randn_64x256"] = (torch.randn(64, 256, generator=g) * 0.02, [7])
    cases["randn_128x1024"] = (torch.randn(128, 1024, generator=g) * 0.02, [7, 500, 1023])
    w = torch.randn(32, 128, generator=g) * 0.02
    w[0, :16] = 0.25           # xmin == xmax -> (-1, +1) degenerate fix
    w[1, :] = 0.0              # zero row
    w[16:32, 16:32] = 0.01     # 16 rows with identical scale in one group -> degenerate 2nd level
    w[5, 48:64] = 3.0
    w[5, 112:128] = 3.0        # 4-bit pool of row 5 constant
    cases["edges_32x128"] = (w, [])
    cases["randn_16x4096"] = (torch.randn(16, 4096, generator=g) * 0.02, [0, 4095])
    for cname, (w32, dead_cols) in cases.items():
        N, K = w32.shape
        layer = torch.nn.Linear(K, N, bias=False)
        layer.weight.data = w32.to(torch.float16)
        W_in = layer.weight.data.clone()
        gpt = mxqgpt_mod.MXQGPT(layer)
        wr = layerwrapper.WrappedGPT(layer)
        X = torch.randn(3, 24, K, generator=g).to(torch.float16)
        X[:, :, dead_cols] = 0
        sr = []
        for j in range(3):
            gpt.add_batch(X[j], None)
            wr.add_batch(X[j], None)
            sr.append(wr.scaler_row.clone().numpy())
        diagH = torch.diag(gpt.H).clone().numpy()
        gpt.fasterquant(percdamp=0.01, blocksize=16)
        key = f"{cname}"
        out[key + "/W"] = W_in.numpy()
        out[key + "/X"] = X.numpy()
        out[key + "/diagH"] = diagH
        out[key + "/scaler_row"] = np.stack(sr)
        out[key + "/Wq"] = layer.weight.data.numpy()
        out[key + "/wanda"] = (torch.abs(W_in.float()) * torch.sqrt(wr.scaler_row.reshape((1, -1)))).numpy()
    np.savez_compressed(os.path.join(OUT, "fasterquant.npz"), **out)
    print("fasterquant.npz", len(out), "arrays")


def gen_fasterquant_blocksize(mxqgpt_mod):
    """fasterquant with the signature's default blocksize=128 (one 48-wide 2-bit group per block,
    mxqgpt.py:388,413-415) and with 32 (groups of 32 + 16)."""
    torch.cuda.synchronize = lambda *a, **k: None
    out = {}
    g = torch.Generator().manual_seed(12)
    w = torch.randn(64, 512, generator=g) * 0.02
    w[3, :48] = 0.125          # constant 48-wide group
    w[16:32, 64:112] = 0.01    # 16 rows sharing one scale
    for bs in (128, 48, 32):
        layer = torch.nn.Linear(512, 64, bias=False)
        layer.weight.data = w.to(torch.float16)
        gpt = mxqgpt_mod.MXQGPT(layer)
        X = torch.randn(2, 24, 512, generator=g).to(torch.float16)
        X[:, :, [9, 300]] = 0
        for j in range(2):
            gpt.add_batch(X[j], None)
        gpt.fasterquant(blocksize=bs)
        out[f"bs{bs}/W"] = w.to(torch.float16).numpy()
        out[f"bs{bs}/X"] = X.numpy()
        out[f"bs{bs}/Wq"] = layer.weight.data.numpy()
    np.savez_compressed(os.path.join(OUT, "fasterquant_blocksize.npz"), **out)
    print("fasterquant_blocksize.npz", len(out), "arrays")


def gen_quantizer(quantizer):
    """Quantizer used directly (a-7): bits 2/4, perchannel, asym, qq_scale_bits=4."""
    out = {}
    g = torch.Generator().manual_seed(3)
    for bits, shape in ((2, (32, 16)), (4, (48, 64)), (4, (16, 1024))):
        x = torch.randn(*shape, generator=g) * 0.03
        q = quantizer.Quantizer()
        q.configure(bits=bits, perchannel=True, sym=False, qq_scale_bits=4)
        q.find_params(x, weight=True)
        y = q.quantize_dequantize(x)
        codes = q.quantize(x)
        key = f"b{bits}_{shape[0]}x{shape[1]}"
        out[key + "/x"] = x.numpy()
        out[key + "/y"] = y.numpy()
        out[key + "/codes"] = codes.numpy().astype(np.uint8)
        out[key + "/scale"] = q.scale.reshape(-1).numpy()
        out[key + "/zero"] = q.zero.reshape(-1).numpy()
    np.savez_compressed(os.path.join(OUT, "quantizer.npz"), **out)
    print("quantizer.npz", len(out), "arrays")


def actquant_cases():
    """name -> (fp32 tensor, layerwise).  Shapes chosen to reach every branch of
    utils_quant.py:50-81 / :130-187, including the 3-D token-slicing quirk (T > (C // G) * G)."""
    g = torch.Generator().manual_seed(4)
    sym, asym = {}, {}
    x = torch.randn(16, 256, generator=g)
    x[2, :] = 0.0                       # zero row: max_input = 0 -> s = 127 / 1e-6
    x[3, :128] = 1e-7                   # max_input comparable to the 1e-6 guard
    x[4, 5] = 2.5
    x[4, 6] = -2.0                      # STE clip plants
    x[5, :] *= 300.0
    sym["2d_16x256"] = (x, False)
    sym["3d_2x12x256"] = (torch.randn(2, 12, 256, generator=g), False)
    sym["3d_dead_1x130x128"] = (torch.randn(1, 130, 128, generator=g), False)   # tokens 128,129 dead
    sym["3d_1x160x1024"] = (torch.randn(1, 160, 1024, generator=g) * 3.0, False)
    sym["4d_2x3x8x16"] = (torch.randn(2, 3, 8, 16, generator=g), False)
    sym["layerwise_4x256"] = (torch.randn(4, 256, generator=g), True)
    y = torch.randn(16, 64, generator=g) * 0.5
    y[0, :8] = 0.25                     # constant group: alpha = 0 -> divide by 1e-8
    y[1, :] = 0.0
    y[2, 3] = 2.5
    y[2, 4] = -2.0
    y[3, :] *= 1e-4
    asym["2d_16x64"] = (y, False)
    asym["3d_dead_2x20x16"] = (torch.randn(2, 20, 16, generator=g), False)      # tokens 16..19 dead
    asym["3d_1x160x1024"] = (torch.randn(1, 160, 1024, generator=g), False)
    asym["4d_2x4x16x8"] = (torch.randn(2, 4, 16, 8, generator=g), False)
    asym["layerwise_4x256"] = (torch.randn(4, 256, generator=g), True)
    return sym, asym


def gen_actquant(utils_quant):
    """SymQuantizer / AsymQuantizer forward + backward (utils_quant.py:31-199)."""
    out = {}
    clip = torch.tensor([-2.0, 2.0])
    sym, asym = actquant_cases()
    for mode, fn, cases in (("sym", utils_quant.SymQuantizer, sym), ("asym", utils_quant.AsymQuantizer, asym)):
        for dname, td in TD.items():
            for cname, (x32, layerwise) in cases.items():
                for bits in (8, 4):
                    if bits == 4 and not cname.startswith(("2d", "3d_dead")):
                        continue
                    x = x32.to(td).clone().requires_grad_(True)
                    y = fn.apply(x, clip, bits, layerwise)
                    key = f"{mode}/{dname}/{cname}/b{bits}"
                    out[key + "/y"] = y.detach().float().numpy()
                    if bits == 8:
                        out[f"{mode}/{dname}/{cname}/x"] = x.detach().float().numpy()
                    if cname.startswith("2d") and bits == 8:
                        gg = torch.Generator().manual_seed(5)
                        go = torch.randn(x.shape, generator=gg).to(td)
                        y.backward(go)
                        out[key + "/go"] = go.float().numpy()
                        out[key + "/gi"] = x.grad.float().numpy()
    np.savez_compressed(os.path.join(OUT, "actquant.npz"), **out)
    print("actquant.npz", len(out), "arrays")


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    torch.manual_seed(0)
    torch.set_num_threads(4)
    utils_quant, mxqgpt_mod, layerwrapper, quantizer = _import_reference()
    gen_fakequant(utils_quant)
    gen_fasterquant(mxqgpt_mod, layerwrapper)
    gen_fasterquant_blocksize(mxqgpt_mod)
    gen_quantizer(quantizer)
    gen_actquant(utils_quant)


if __name__ == "__main__":
    main()
