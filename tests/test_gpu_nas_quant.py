"""A live run of the `--prune_method mxq` driver (mxq_quant/lib/prune.py:326-425, main.py:29-100)
on a tiny random-init Llama: calibration capture through the Catcher, forward hooks feeding
MXQGPT.add_batch during the per-sample layer forwards, fasterquant, re-forward, in/out swap.
Every decoder linear must end up bit-identical to the oracle's fasterquant of its ORIGINAL weight
with the dead columns the calibration really produced, and `args.pack` must attach packed tensors
that decode to the oracle packer's result."""
import argparse
import copy

import numpy as np
import pytest
import torch

from oracle import mxq_oracle as O
from tests.gpu_util import packed_to_np

pytestmark = pytest.mark.gpu

HIDDEN, INTER, LAYERS, SEQ, NSAMPLES, VOCAB = 256, 704, 2, 64, 4, 128
DEAD_ATTN, DEAD_MLP = 7, 5


def tiny_llama(dev):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=HIDDEN, intermediate_size=INTER, num_hidden_layers=LAYERS,
                      num_attention_heads=4, num_key_value_heads=4, vocab_size=VOCAB,
                      max_position_embeddings=SEQ, rms_norm_eps=1e-5, use_cache=False)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).half().to(dev).eval()
    model.seqlen = SEQ                                   # main.py:26
    with torch.no_grad():
        for layer in model.model.layers:
            # an exactly-zero RMSNorm gain makes that input channel of q/k/v (gate/up) zero for every
            # token: a "dead" column (mxqgpt.py:399-403) that is known without re-running the model
            layer.input_layernorm.weight[DEAD_ATTN] = 0
            layer.post_attention_layernorm.weight[DEAD_MLP] = 0
    return model


def calib(dev):
    g = torch.Generator().manual_seed(1)
    return [(torch.randint(0, VOCAB, (1, SEQ), generator=g).to(dev), None) for _ in range(NSAMPLES)]


def expected_dead(name, K):
    dead = np.zeros(K, bool)
    if name.split(".")[-1] in ("q_proj", "k_proj", "v_proj"):
        dead[DEAD_ATTN] = True
    if name.split(".")[-1] in ("gate_proj", "up_proj"):
        dead[DEAD_MLP] = True
    return dead


def check_model(model, originals, packed_required):
    from mxq_b200.prune import find_layers
    n = 0
    for li, layer in enumerate(model.model.layers):
        for name, lin in find_layers(layer).items():
            W0 = originals[(li, name)]
            dead = expected_dead(name, W0.shape[1])
            want = O.fasterquant(W0, dead)
            got = lin.weight.data.cpu().numpy()
            assert got.dtype == np.float16 and got.shape == W0.shape
            assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), f"layer {li} {name}"
            if packed_required:
                assert hasattr(lin, "mxq_packed"), f"layer {li} {name}: no packed tensors attached"
                wantp = O.pack_mxq(W0, dead)
                gotp = packed_to_np(lin.mxq_packed)
                for k in wantp:
                    a = gotp[k].view(np.uint16) if gotp[k].dtype == np.float16 else gotp[k]
                    b = wantp[k].view(np.uint16) if wantp[k].dtype == np.float16 else wantp[k]
                    assert np.array_equal(a, b), f"layer {li} {name} packed {k}"
                assert np.array_equal(O.decode_mxq(gotp), O.decode_mxq(wantp))
            n += 1
    assert n == LAYERS * 7


def snapshot(model):
    from mxq_b200.prune import find_layers
    return {(li, name): lin.weight.data.cpu().numpy().copy()
            for li, layer in enumerate(model.model.layers) for name, lin in find_layers(layer).items()}


@pytest.mark.parametrize("pack", [False, True])
def test_nas_quant_live(cuda, pack, capsys):
    from mxq_b200 import prune
    model = tiny_llama(cuda)
    originals = snapshot(model)
    ref_model = copy.deepcopy(model)
    args = argparse.Namespace(nsamples=NSAMPLES, seed=0, save=None, pack=pack)
    prune.nas_quant(args, model, None, cuda, dataloader=calib(cuda))
    assert model.config.use_cache is False
    check_model(model, originals, pack)
    out = capsys.readouterr().out
    assert "Starting ..." in out and "Ready." in out and out.count("Pruning ...") == LAYERS * 7   # prune.py:327,366,408
    # the quantized model still runs and differs from the fp16 one (weights really were replaced)
    ids = calib(cuda)[0][0]
    with torch.no_grad():
        a = model(ids).logits.float()
        b = ref_model(ids).logits.float()
    assert torch.isfinite(a).all() and not torch.equal(a, b)


def test_nas_quant_early_exit_and_batched_capture_keep_the_statistics(cuda, monkeypatch):
    """The two extensions that save forward time -- stopping a layer's FIRST forward once its last linear has
    seen its input, and capturing / forwarding several calibration samples per call -- must leave every
    linear's calibration statistic (and therefore the quantized model) unchanged."""
    from mxq_b200 import prune
    from mxq_b200.mxqgpt import MXQGPT
    base = tiny_llama(cuda)

    def run(early_exit, batch_size):
        model = copy.deepcopy(base)
        stats = []
        orig = MXQGPT.fasterquant

        def spy(self, *a, **k):
            stats.append((self.nsamples, self.diagH.clone()))
            return orig(self, *a, **k)
        monkeypatch.setattr(MXQGPT, "fasterquant", spy)
        args = argparse.Namespace(nsamples=NSAMPLES, seed=0, save=None, pack=False)
        prune.nas_quant(args, model, None, cuda, dataloader=calib(cuda), batch_size=batch_size, early_exit=early_exit)
        monkeypatch.setattr(MXQGPT, "fasterquant", orig)
        return stats, snapshot(model)

    ref_stats, ref_w = run(False, 1)
    assert len(ref_stats) == LAYERS * 7
    for early_exit, bsz in ((True, 1), (True, NSAMPLES), (False, NSAMPLES)):
        stats, w = run(early_exit, bsz)
        assert len(stats) == len(ref_stats)
        for (n0, d0), (n1, d1) in zip(ref_stats, stats):
            assert n0 == n1 == NSAMPLES
            if bsz == 1:
                assert torch.equal(d0, d1)                                   # same calls, same order
            else:
                assert torch.allclose(d0, d1, rtol=2e-3, atol=0) and torch.equal(d0 == 0, d1 == 0)
        if bsz == 1:
            for k in ref_w:
                assert np.array_equal(ref_w[k].view(np.uint16), w[k].view(np.uint16)), k


def test_nas_quant_layer_callback_sees_final_weights(cuda):
    """layer_callback(i, layer) fires once per layer, after its linears were replaced and (args.pack) annotated,
    before the layer's second forward: what a caller needs to stream results out under the remaining forwards."""
    from mxq_b200 import prune
    model = tiny_llama(cuda)
    originals = snapshot(model)
    seen = []

    def cb(i, layer):
        lins = prune.find_layers(layer)
        assert all(hasattr(m, "mxq_packed") for m in lins.values())
        seen.append((i, {n: m.weight.data.clone() for n, m in lins.items()}))

    args = argparse.Namespace(nsamples=NSAMPLES, seed=0, save=None, pack=True)
    prune.nas_quant(args, model, None, cuda, dataloader=calib(cuda), batch_size=2, layer_callback=cb)
    assert [i for i, _ in seen] == list(range(LAYERS))
    for i, ws in seen:
        for n, w in ws.items():
            assert torch.equal(w, prune.find_layers(model.model.layers[i])[n].weight.data)
            assert not np.array_equal(w.cpu().numpy(), originals[(i, n)])
    check_model(model, originals, True)


def test_nas_quant_needs_dataloader_offline(cuda):
    from mxq_b200 import prune
    model = tiny_llama(cuda)
    with pytest.raises(RuntimeError, match="dataloader"):
        prune.nas_quant(argparse.Namespace(nsamples=2, seed=0, save=None), model, None, cuda)


def test_main_cli_synthetic_calibration(cuda, tmp_path):
    """python -m mxq_b200.main --model DIR --prune_method mxq --synthetic_calib ... (main.py:29-100)."""
    from mxq_b200 import main as cli
    model = tiny_llama(cuda)
    originals = snapshot(model)
    src, dst = tmp_path / "src", tmp_path / "dst"
    model.save_pretrained(src)
    cli.main(["--model", str(src), "--prune_method", "mxq", "--nsamples", str(NSAMPLES), "--seqlen", str(SEQ),
              "--synthetic_calib", "--save_model", str(dst), "--pack"])
    from transformers import AutoModelForCausalLM
    reloaded = AutoModelForCausalLM.from_pretrained(dst, torch_dtype=torch.float16).to(cuda)
    check_model(reloaded, originals, False)
    # the packed checkpoint written next to it loads into MXQLinear modules that reproduce the decode
    from mxq_b200.packed_linear import MXQLinear, load_packed
    load_packed(reloaded, str(dst / "mxq_packed.pt"))
    lin = reloaded.model.layers[0].self_attn.q_proj
    assert isinstance(lin, MXQLinear)
    W0 = originals[(0, "self_attn.q_proj")]
    want = O.decode_mxq(O.pack_mxq(W0, expected_dead("self_attn.q_proj", HIDDEN)))
    assert np.array_equal(lin.dequantize(torch.float32).cpu().numpy(), want)
