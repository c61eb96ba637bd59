"""Parameter containers that reproduce the reference's ``state_dict`` layout key for key.

The reference's checkpoint keys come from its module nesting (``decoder.transformer.h.0.attn.c_attn.weight`` ...,
SURVEY.md Appendix B).  The B200 path does not need that nesting to compute -- every kernel takes raw pointers -- so
the modules here are plain containers built from ``model_spec.state_schema``: same names, shapes, dtypes, tying and
parameter/buffer split, therefore ``state_dict()`` / ``load_state_dict()`` / ``named_parameters()`` and the reference's
``PatternMatcher`` globs keep working unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

BUFFER_LEAVES = ("projection_mat", "grid", "pos_offset")      # reference models/layers.py:119-134 register_buffer
FROZEN_KEYS = ("peer_proj_wt",)                               # reference models/encoder.py:92-95 requires_grad=False


class ParamTree(nn.Module):
    """An nn.Module whose children/parameters are created from dotted keys."""

    def add(self, dotted: str, tensor: torch.Tensor, param: Optional[nn.Parameter] = None) -> nn.Parameter:
        parts = dotted.split(".")
        node = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamTree())
            node = node._modules[p]
        leaf = parts[-1]
        if leaf in BUFFER_LEAVES:
            node.register_buffer(leaf, tensor, persistent=True)
            return None
        if param is None:
            # (peer_proj_wt is a frozen (1,) dummy unless the PEER tail is configured: reference models/encoder.py:86-95)
            frozen = any(dotted.endswith(k) for k in FROZEN_KEYS) and tensor.numel() == 1
            param = nn.Parameter(tensor, requires_grad=not frozen)
        node.register_parameter(leaf, param)
        return param


def reference_like_init(key: str, shape, dtype, spec: dict, gen: Optional[torch.Generator]) -> torch.Tensor:
    """Initial values with the distributions the reference's constructors use (not its RNG stream):
    models/decoder.py:193-212 (N(0,0.02) linears/embeddings, zero biases, c_proj / sqrt(2 n_layer)),
    torchvision ViT defaults, nn.EmbeddingBag N(0,1), unit-column LSH projections (models/layers.py:119-134)."""
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("peer_proj_wt"):
        return torch.zeros(shape) if len(shape) == 1 else torch.randn(shape, generator=gen) / math.sqrt(shape[0])
    if ".peer." in key:                        # nn.Embedding N(0, 1); nn.Linear ~ U(+-1/sqrt(in)) (same scale, not the same stream)
        return torch.randn(shape, generator=gen) * (1.0 if ".emb_" in key else 1.0 / math.sqrt(3 * shape[-1]))
    if leaf == "projection_mat":
        return torch.nn.functional.normalize(torch.randn(shape, generator=gen), p=2.0, dim=0)
    if leaf == "grid":
        nb = shape[0]
        return torch.linspace(-1, 1, nb + 1)[:-1] + 0.5 * (2.0 / nb)
    if leaf == "pos_offset":
        nb1 = spec["lsh_num_bins"][int(key.split(".emb.")[1].split(".")[0])] + 1
        return (nb1 * torch.arange(0, shape[0], dtype=torch.long)).reshape(shape)
    if ".lsh_emb." in key and key.endswith("emb.weight"):
        return torch.randn(shape, generator=gen)
    if len(shape) == 1:
        is_norm_gain = leaf == "weight"
        return torch.ones(shape) if is_norm_gain else torch.zeros(shape)
    if leaf in ("class_token",):
        return torch.zeros(shape)
    std = 0.02
    if key.startswith("decoder.") and key.endswith("c_proj.weight"):
        std = 0.02 / math.sqrt(2 * spec["n_layer"])
    if ".proj.models." in key or key == "encoder.1.weight":
        std = 1.0 / math.sqrt(shape[-1])
    return (std * torch.randn(shape, generator=gen)).to(dtype)


def build_param_tree(root: nn.Module, schema, spec: dict, tied, device, gen: Optional[torch.Generator] = None):
    """Populate ``root`` (which must expose ParamTree children ``encoder`` / ``decoder``) from the schema."""
    made: Dict[str, nn.Parameter] = {}
    tie_to = {a: b for a, b in tied}
    # the reference registers `transformer` before `lm_head`, so named_parameters() reports the wte name
    for key, (shape, dtype) in schema.items():
        if key in tie_to:
            continue
        top, rest = key.split(".", 1)
        t = reference_like_init(key, shape, dtype, spec, gen).to(dtype).to(device)
        p = getattr(root, top).add(rest, t)
        if p is not None:
            made[key] = p
    for a, b in tied:
        if a in schema and b in made:
            top, rest = a.split(".", 1)
            getattr(root, top).add(rest, None, param=made[b])
