"""Synthetic inputs of the shapes the reference trains and captions on.

Flickr30K and the hub tokenizer are unreachable offline, so images are N(0,1) tensors (the
statistics of the ViT preprocessing output, reference ``trainer.py:73``) and captions are random
token ids of a Flickr-like length, padded with the ignore index exactly the way
``training/utils.py:16-20`` (``normalize_label``) leaves them: caption tokens, ONE trailing EOS,
then ``-100``.
"""
import torch


def synth_images(n: int, size: int = 224, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(n, 3, size, size, generator=g)


def synth_labels(n: int, width: int = 256, vocab_size: int = 50257, seed: int = 1234, min_len: int = 8,
                 max_len: int = 40, eos: int = 50256, ignore_index: int = -100) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    max_len = min(max_len, width - 1)
    lens = torch.randint(min_len, max_len + 1, (n,), generator=g)
    toks = torch.randint(0, vocab_size, (n, width), generator=g)
    pos = torch.arange(width).unsqueeze(0)
    labels = torch.where(pos < lens.unsqueeze(1), toks, torch.full_like(toks, ignore_index))
    labels[torch.arange(n), lens] = eos
    return labels
