"""CPU: packed-layout oracle -- the reference's known-answer vector, bit positions, and the
encode -> decode round trip of our packer policy."""
import numpy as np
import pytest

from oracle import mxq_oracle as O


def test_reference_kat_constant_fill():
    # cuda_kernel/test_correct_gemv.py:19-53: every weight decodes to 1 -> y == 4096
    p = O.kat_constant_fill(64, 4096)
    assert (O.decode_mxq(p) == 1).all()
    y = O.gemv_mxq(np.ones((1, 4096), np.float16), p)
    assert (y == 4096).all()


def test_packed_shapes_reference_sizes():
    shp = O.packed_shapes(4096, 4096)
    assert shp["weight"] == (4096, 256) and shp["weight_last"] == (4096, 64)
    assert shp["zeros_and_scales"] == (4096, 32) and shp["zeros_2nd"] == (1024, 32)
    assert shp["scales_2nd"] == (1024, 192) and shp["zeros_4b"] == (512,)
    nbytes = sum(int(np.prod(s)) * (2 if k.startswith("scales") else 4) for k, s in shp.items())
    assert nbytes == 6301696
    assert O.packed_shapes(4096, 11008)["zeros_and_scales"] == (4096, 96)
    with pytest.raises(ValueError):
        O.packed_shapes(4096, 100)


def test_bit_positions_single_code():
    # one 2-bit code and one 4-bit code set, everything else zero: pins bit/column order
    OC, IC = 8, 128
    shp = O.packed_shapes(OC, IC)
    p = {k: np.zeros(s, dtype=np.float16 if k.startswith("scales") else np.int32) for k, s in shp.items()}
    p["scales_2nd"][:] = 1
    p["scales_4b"][:] = 1
    # scale code c=1 for every group, zeros 0  -> weight = q
    p["zeros_and_scales"][:] = np.uint32(0x15001500).view(np.int32)  # cbyte=0b010101 both halves
    # row 3, block 1, slot 2, j=5 (col 64+32+5=101) code 3
    p["weight"][3, 4 * 1 + 2] = 3 << (2 * 5)
    # row 2, block 0, col 48+6 nibble 9 ; row 2, block 1, col 64+56+7 nibble 12
    p["weight"][2, 3] = 9 << (4 * 6)
    p["weight_last"][2, 1] = np.uint32(12 << (4 * 7)).view(np.int32)
    W = O.decode_mxq(p)
    want = np.zeros((OC, IC), np.float32)
    want[3, 101] = 3
    want[2, 54] = 9
    want[2, 127] = 12
    assert np.array_equal(W, want)


def test_metadata_halfword_positions():
    OC, IC = 8, 64 * 70      # 70 blocks: second chunk of metadata words in use
    shp = O.packed_shapes(OC, IC)
    assert shp["zeros_and_scales"] == (OC, 64)
    p = {k: np.zeros(s, dtype=np.float16 if k.startswith("scales") else np.int32) for k, s in shp.items()}
    p["scales_2nd"][:] = 2
    p["weight"][:] = np.uint32(0xFFFFFFFF).view(np.int32)   # q = 3 everywhere (2-bit) / 15 (4-bit)
    # block 37 -> word 5, half 1; block 66 -> chunk 1, word 32+2, half 0
    p["zeros_and_scales"][1, 5] = np.uint32((0b100110 << 8 | 0b000001) << 16).view(np.int32)
    p["zeros_and_scales"][1, 34] = (0b11 << 8 | 0b10)
    p["zeros_2nd"][0, 34] = 0b01
    W = O.decode_mxq(p).reshape(OC, 70, 64)
    # block 37: c = (2,1,2), z1 = (1,0,0), z2 = 0 -> w = 2*c*(3 - z1)
    assert W[1, 37, 0] == 2 * 2 * 2 and W[1, 37, 16] == 2 * 1 * 3 and W[1, 37, 32] == 2 * 2 * 3
    # block 66 slot 0: c=3, z1=2, z2=1 -> 2*(3-1)*(3-2) = 4
    assert W[1, 66, 0] == 4
    assert W[1, 36, 0] == 0 and W[2, 37, 0] == 0


@pytest.mark.parametrize("shape", [(16, 64), (32, 256), (16, 4096 + 128)])
def test_pack_roundtrip_error_bound(shape):
    rng = np.random.default_rng(1)
    W = (rng.standard_normal(shape) * 0.02).astype(np.float16)
    p = O.pack_mxq(W)
    for k, s in O.packed_shapes(*shape).items():
        assert p[k].shape == s, k
    Wd = O.decode_mxq(p)
    err = np.abs(Wd - W.astype(np.float32))
    # 2-bit RTN noise: rms error well under the weight rms, and codes decode inside the range
    assert np.sqrt((err ** 2).mean()) < 0.45 * 0.02
    # idempotence of the codes: packing the decoded weights reproduces the same decode
    Wd2 = O.decode_mxq(O.pack_mxq(Wd.astype(np.float16)))
    assert np.sqrt(((Wd2 - Wd) ** 2).mean()) < 0.2 * 0.02


def test_pack_dead_columns_and_zero_rows():
    rng = np.random.default_rng(2)
    W = (rng.standard_normal((16, 128)) * 0.02).astype(np.float16)
    W[5] = 0
    dead = np.zeros(128, bool)
    dead[[3, 50, 127]] = True
    Wd = O.decode_mxq(O.pack_mxq(W, dead))
    assert (Wd[:, dead] == 0).all()       # zero is always representable (integer zero-points)
    assert (Wd[5] == 0).all()


def test_random_packed_decode_matches_scalar_formula():
    p = O.random_packed(8, 128, seed=3)
    W = O.decode_mxq(p)
    wt = p["weight"].view(np.uint32); wl = p["weight_last"].view(np.uint32)
    zs = p["zeros_and_scales"].view(np.uint32); z2 = p["zeros_2nd"].view(np.uint32)
    for oc, col in [(0, 0), (3, 17), (7, 47), (5, 64 + 33), (2, 50), (6, 64 + 60)]:
        blk, r = divmod(col, 64)
        if r < 48:
            k, j = divmod(r, 16)
            q = (int(wt[oc, 4 * blk + k]) >> (2 * j)) & 3
            hw = (int(zs[oc, blk % 32]) >> (16 * (blk // 32))) & 0xFFFF
            z1 = (hw >> (2 * k)) & 3
            c = (hw >> (8 + 2 * k)) & 3
            zz = ((int(z2[oc // 4, blk % 32]) >> (8 * (blk // 32))) >> (2 * k)) & 3
            want = np.float32(np.float32(p["scales_2nd"][oc // 4, 3 * blk + k]) * np.float32(c - zz)) * np.float32(q - z1)
        else:
            j = r - 48
            word = int(wt[oc, 4 * blk + 3]) if j < 8 else int(wl[oc, blk])
            q = (word >> (4 * (j % 8))) & 0xF
            z4 = (int(p["zeros_4b"].view(np.uint32)[oc // 8]) >> (4 * (oc % 8))) & 0xF
            want = np.float32(p["scales_4b"][oc]) * np.float32(q - z4)
        assert W[oc, col] == np.float32(want)


def test_allocate_group_bits_oracle():
    """f-3: the most important of every 4 groups gets the pooled 4-bit slot; ties -> lowest index;
    with uniform importance the mask differs from the positional recipe only in which slot is 4-bit."""
    rng = np.random.default_rng(0)
    W = (rng.standard_normal((32, 256)) * 0.02).astype(np.float16)
    W[:, 16:32] *= 8                       # block 0: group 1 salient
    W[:, 64 + 32:64 + 48] *= 8             # block 1: group 2 salient
    gb, imp = O.allocate_group_bits(W)
    assert gb.reshape(-1, 4)[0].tolist() == [2, O.POOL | 4, 2, 2]
    assert gb.reshape(-1, 4)[1].tolist() == [2, 2, O.POOL | 4, 2]
    assert (gb == (O.POOL | 4)).sum() == 256 // 64 and imp.shape == (16,)
    # activation statistics can override the weight magnitudes
    sr = np.ones(256, np.float32)
    sr[0:16] = 1e4
    gb2, _ = O.allocate_group_bits(W, sr)
    assert gb2.reshape(-1, 4)[0].tolist() == [O.POOL | 4, 2, 2, 2]
    # ties
    Wc = np.full((16, 64), 0.5, np.float16)
    assert O.allocate_group_bits(Wc)[0].tolist() == [O.POOL | 4, 2, 2, 2]
    # the mask is a valid recipe for the fake quantizer and lowers the error on the salient columns
    x = W.astype(np.float32)
    e_pos = np.abs(O.fakequant_fwd(x, "fp32", 2) - x)[:, 16:32].mean()
    e_imp = np.abs(O.fakequant_fwd(x, "fp32", 2, group_bits=gb) - x)[:, 16:32].mean()
    assert e_imp < e_pos
