"""CLI mirror of mxq_quant/main.py:29-100 for the ``--prune_method mxq`` entry.

    python -m mxq_b200.main --model <hf-path> --prune_method mxq [--nsamples 128] [--save_model DIR]
                            [--pack] [--synthetic_calib [--seqlen 2048]]

The reference downloads wikitext2 for calibration (lib/data.py, network); offline,
``--synthetic_calib`` (an extension) draws `nsamples` random token sequences instead, seeded by
``--seed``.  ``--pack`` (extension) also produces the packed 2/4-bit tensors and, with
``--save_model``, writes them as ``mxq_packed.pt`` next to the fp16 fake-quant checkpoint the
reference saves (main.py:96-100).
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch


def get_llm(model, seqlen=2048):
    from transformers import AutoModelForCausalLM
    m = AutoModelForCausalLM.from_pretrained(model, torch_dtype=torch.float16, low_cpu_mem_usage=True)
    m.seqlen = seqlen                                            # main.py:26
    return m.cuda()


def synthetic_loader(nsamples, seqlen, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randint(0, vocab, (1, seqlen), generator=g), None) for _ in range(nsamples)]


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--model', type=str, help='LLaMA model')
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--nsamples', type=int, default=128)
    parser.add_argument('--sparsity_ratio', type=float, default=0)
    parser.add_argument("--sparsity_type", type=str, default=None)
    parser.add_argument("--prune_method", type=str, choices=["mxq"], default="mxq")
    parser.add_argument('--save', type=str, default=None)
    parser.add_argument('--save_model', type=str, default=None)
    parser.add_argument('--pack', action="store_true",
                        help="also attach the packed 2/4-bit tensors to every linear (extension)")
    parser.add_argument('--synthetic_calib', action="store_true",
                        help="random-token calibration instead of wikitext2 (extension; offline)")
    parser.add_argument('--seqlen', type=int, default=2048)
    args = parser.parse_args(argv)
    if args.sparsity_ratio != 0:
        raise SystemExit("only --prune_method mxq with --sparsity_ratio 0 is implemented")
    np.random.seed(args.seed)
    torch.random.manual_seed(args.seed)
    from .prune import nas_quant
    model = get_llm(args.model, args.seqlen)
    model.eval()
    tokenizer = None
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(args.model, use_fast=False)
    except Exception:
        if not args.synthetic_calib:
            raise
    loader = synthetic_loader(args.nsamples, args.seqlen, model.config.vocab_size, args.seed) \
        if args.synthetic_calib else None
    nas_quant(args, model, tokenizer, torch.device("cuda:0"), dataloader=loader)
    if args.save_model:
        if args.pack:
            os.makedirs(args.save_model, exist_ok=True)
            linears = {n: {k: v.detach().cpu() for k, v in m.mxq_packed.items()}
                       for n, m in model.named_modules() if hasattr(m, "mxq_packed")}
            from .packed_linear import FORMAT
            torch.save({"format": FORMAT, "linears": linears}, os.path.join(args.save_model, "mxq_packed.pt"))
        model.save_pretrained(args.save_model)
        if tokenizer is not None:
            tokenizer.save_pretrained(args.save_model)


if __name__ == '__main__':
    main()
