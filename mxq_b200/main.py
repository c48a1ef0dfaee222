"""CLI mirror of mxq_quant/main.py:29-100 for the ``--prune_method mxq`` entry.

    python -m mxq_b200.main --model <hf-path> --prune_method mxq [--nsamples 128] [--save_model DIR] [--pack]

Needs a local Hugging Face checkpoint and a calibration set (the reference downloads wikitext2);
offline, drive ``mxq_b200.prune.nas_quant(..., dataloader=...)`` or ``quantize_linear`` directly.
"""
from __future__ import annotations

import argparse

import numpy as np
import torch


def get_llm(model):
    from transformers import AutoModelForCausalLM
    m = AutoModelForCausalLM.from_pretrained(model, torch_dtype=torch.float16, low_cpu_mem_usage=True)
    m.seqlen = 2048                                              # main.py:26
    return m.cuda()


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
    args = parser.parse_args(argv)
    if args.sparsity_ratio != 0:
        raise SystemExit("only --prune_method mxq with --sparsity_ratio 0 is implemented")
    np.random.seed(args.seed)
    torch.random.manual_seed(args.seed)
    from transformers import AutoTokenizer
    from .prune import nas_quant
    model = get_llm(args.model)
    model.eval()
    tokenizer = AutoTokenizer.from_pretrained(args.model, use_fast=False)
    nas_quant(args, model, tokenizer, torch.device("cuda:0"))
    if args.save_model:
        model.save_pretrained(args.save_model)
        tokenizer.save_pretrained(args.save_model)


if __name__ == '__main__':
    main()
