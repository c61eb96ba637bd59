"""Shared builders for the parity tests (weights are re-synthesised, never stored)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from image2text_b200.config_schema import load_training_config  # noqa: E402
from image2text_b200.model_spec import spec_from_config, synth_state_dict  # noqa: E402

_CACHE = {}

SPEC_OVERRIDES = {"tiny": dict(vit_layers=2, vit_image=32), "nano": {}, "gpt2": {}}


def spec_and_weights(name: str, seed: int = 0):
    key = (name, seed)
    if key not in _CACHE:
        tc = load_training_config(os.path.join(ROOT, "configs", name + ".yaml"))
        spec = spec_from_config(tc.model, **SPEC_OVERRIDES[name])
        _CACHE[key] = (tc, spec, synth_state_dict(spec, seed=seed))
    return _CACHE[key]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (the '1e-4 relative' of BASELINE.json is read as relative to the tensor scale)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
