"""CPU: the numpy oracle is pinned bit-for-bit to vectors produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest

from oracle import mxq_oracle as O


def _cases(npz, depth):
    return sorted({"/".join(k.split("/")[:depth]) for k in npz.files})


@pytest.fixture(scope="module")
def fq(golden_dir):
    return np.load(os.path.join(golden_dir, "fakequant.npz"))


@pytest.fixture(scope="module")
def fqt(golden_dir):
    return np.load(os.path.join(golden_dir, "fasterquant.npz"))


def test_fakequant_fwd_bit_exact(fq):
    n = 0
    for key in _cases(fq, 2):
        dtype, case = key.split("/")
        nb = 3 if "bits3" in case else 4 if "bits4" in case else 2
        x, y = fq[key + "/x"], fq[key + "/y"]
        out = O.fakequant_fwd(x, dtype, nb)
        assert np.array_equal(out.view(np.uint32), y.view(np.uint32)), key
        n += 1
    assert n >= 17


def test_ste_bwd_bit_exact(fq):
    for key in _cases(fq, 2):
        if key + "/gi" not in fq.files:
            continue
        gi = O.ste_bwd(fq[key + "/go"], fq[key + "/x"])
        assert np.array_equal(gi.view(np.uint32), fq[key + "/gi"].view(np.uint32)), key


def test_ste_boundary_inclusive():
    x = np.array([[2.0, -2.0, 1.9999999, -1.9999999, 2.5, 0.0]], dtype=np.float32)
    g = np.ones_like(x)
    assert O.ste_bwd(g, x).tolist() == [[0, 0, 1, 1, 0, 1]]


def test_survey_sin_codes(fq):
    # SURVEY.md 8c self-check numbers produced from the reference
    x = fq["fp32/sin_2x128/x"]
    _, q, *_ = O.fakequant_fwd(x, "fp32", 2, return_aux=True)
    want = [1, 2, 3, 3, 3, 3, 3, 2, 2, 1, 1, 0, 0, 0, 0, 0, 1, 2, 2, 3, 3, 3, 3, 3, 2, 2, 1, 1, 0, 0, 0, 0,
            0, 1, 2, 2, 3, 3, 3, 3, 3, 2, 2, 1, 1, 0, 0, 0, 1, 2, 5, 8, 10, 13, 14, 15, 15, 13, 11, 9, 6, 3, 1, 0]
    assert q[0, :64].astype(int).tolist() == want


def test_fasterquant_bit_exact(fqt):
    for key in _cases(fqt, 1):
        W, X, Wq = fqt[key + "/W"], fqt[key + "/X"], fqt[key + "/Wq"]
        dead = O.dead_columns(X)
        assert np.array_equal(dead, fqt[key + "/diagH"] == 0), key
        out = O.fasterquant(W, dead)
        assert np.array_equal(out.view(np.uint16), Wq.view(np.uint16)), key


def test_scaler_row_and_wanda(fqt):
    for key in _cases(fqt, 1):
        W, X = fqt[key + "/W"], fqt[key + "/X"]
        sr, n = np.zeros(W.shape[1], np.float32), 0
        for j in range(X.shape[0]):
            sr, n = O.scaler_row_update(sr, n, X[j])
            ref = fqt[key + "/scaler_row"][j]
            # torch.norm()**2 sums in a different order: tolerance 1e-6 relative
            assert np.abs(sr - ref).max() <= 1e-6 * np.abs(ref).max()
        wm = O.wanda_metric(W, sr)
        assert np.abs(wm - fqt[key + "/wanda"]).max() <= 1e-6 * np.abs(wm).max()
        # diag(H) == 2 * scaler_row (SURVEY a-9 note)
        assert np.allclose(fqt[key + "/diagH"], 2 * sr, rtol=1e-5, atol=0)


def test_quantizer_direct(golden_dir):
    Q = np.load(os.path.join(golden_dir, "quantizer.npz"))
    for key in _cases(Q, 1):
        x = Q[key + "/x"]
        bits = int(key[1])
        mq = 2 ** bits - 1
        scale, zero = O._find_params(x.min(1), x.max(1), mq)
        sq, _, _, _ = O._qq_scale(scale)
        y, q = O._quant_dequant(x, sq[:, None], zero[:, None], mq)
        assert np.array_equal(y, Q[key + "/y"])
        assert np.array_equal(q.astype(np.uint8), Q[key + "/codes"])
        assert np.array_equal(sq, Q[key + "/scale"]) and np.array_equal(zero, Q[key + "/zero"])


def test_actquant_bit_exact(golden_dir):
    """SymQuantizer / AsymQuantizer restatements vs the unmodified reference (all ranks, layerwise,
    the 3-D token-slicing quirk, fp32 / bf16 / fp16; NaNs of 0/0 groups must coincide)."""
    A = np.load(os.path.join(golden_dir, "actquant.npz"))
    n = 0
    for k in sorted(k for k in A.files if k.endswith("/y")):
        mode, dtype, case, b, _ = k.split("/")
        x = A[f"{mode}/{dtype}/{case}/x"]
        fn = O.sym_quant if mode == "sym" else O.asym_quant
        y = fn(x, dtype, int(b[1:]), case.startswith("layerwise"))
        ref = A[k]
        same = (y.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(y) & np.isnan(ref))
        assert same.all(), k
        if k[:-2] + "/gi" in A.files:
            gi = O.ste_bwd(A[k[:-2] + "/go"], x)
            assert np.array_equal(gi.view(np.uint32), A[k[:-2] + "/gi"].view(np.uint32)), k
        n += 1
    assert n >= 45


def test_fasterquant_blocksize_bit_exact(golden_dir):
    """fasterquant(blocksize=128 / 48 / 32) of the unmodified reference (oracle/gen_golden.py)."""
    d = np.load(os.path.join(golden_dir, "fasterquant_blocksize.npz"))
    for bs in (128, 48, 32):
        W, X, Wq = d[f"bs{bs}/W"], d[f"bs{bs}/X"], d[f"bs{bs}/Wq"]
        out = O.fasterquant_blocksize(W, O.dead_columns(X), bs)
        assert np.array_equal(out.view(np.uint16), Wq.view(np.uint16)), bs
    W, X = d["bs128/W"], d["bs128/X"]
    dead = O.dead_columns(X)
    assert np.array_equal(O.fasterquant_blocksize(W, dead, 16).view(np.uint16), O.fasterquant(W, dead).view(np.uint16))
