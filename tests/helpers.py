"""Shared builders for the parity tests (weights are re-synthesised, never stored)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from image2text_b200.config_schema import load_training_config  # noqa: E402
from image2text_b200.model_spec import spec_from_config, synth_state_dict  # noqa: E402

_CACHE = {}

# the reference-made goldens were produced with every dropout at 0 (tests/golden/make_golden.py: zero_dropout / .eval()),
# so the parity models switch the YAML's p = 0.1 off; the dropout tests switch it back on explicitly
NO_DROPOUT = dict(dropout=0.0, attn_dropout=0.0)
SPEC_OVERRIDES = {"tiny": dict(vit_layers=2, vit_image=32), "tiny_peer": dict(vit_layers=2, vit_image=32),
                  "nano": dict(NO_DROPOUT), "gpt2": dict(NO_DROPOUT)}


def spec_and_weights(name: str, seed: int = 0):
    key = (name, seed)
    if key not in _CACHE:
        tc = load_training_config(os.path.join(ROOT, "configs", name + ".yaml"))
        spec = spec_from_config(tc.model, **SPEC_OVERRIDES[name])
        _CACHE[key] = (tc, spec, synth_state_dict(spec, seed=seed))
    return _CACHE[key]


def check_picks_vs_oracle(name, m, images, got, n_prompt_tokens, top_k, tol_frac=2e-2):
    """Teacher-forced parity of a bf16 decode against the fp32 CPU ORACLE (oracle/i2t_oracle.py, pinned to the unmodified
    reference by tests/golden): oracle logits over got[:, :-1]; every pick must be un-banned (no-repeat-n-gram rule) and within
    tol_frac * max|logit| (BASELINE.json's bf16 tolerance) of the oracle's k-th best un-banned logit.  The oracle is given the
    encoder output the model under test computed: the LSH tail is an integer hash of the ViT feature, so a bf16 feature may land
    in a neighbouring bucket (the bf16 ViT trunk itself is pinned by test_gpu_model.py).  Returns (worst gap, logit scale)."""
    from oracle import i2t_oracle as O
    _, spec, sd = spec_and_weights(name)
    with torch.no_grad():
        enc = m.encoder(images).float().cpu()
        _, logits, _ = O.ved_forward(sd, spec, None, got[:, :-1].cpu(), encoder_output=enc, normalize_grads=False)
    logits = logits.float()[..., :spec["vocab_size"]]
    got = got.cpu()
    scale = float(logits.abs().max())
    worst = 0.0
    for t in range(n_prompt_tokens - 1, got.shape[1] - 1):
        row = logits[:, t]
        allowed = O.apply_ngram_ban(got[:, :t + 1], row.clone(), spec["no_repeat_n_grams"])
        pick = got[:, t + 1:t + 2]
        assert bool(torch.isfinite(allowed.gather(1, pick)).all()), f"banned token picked at position {t + 1}"
        kth = torch.topk(allowed, top_k, dim=-1).values[:, -1:]
        gap = (kth - row.gather(1, pick)).clamp_min(0)
        worst = max(worst, float(gap.max()))
        assert float(gap.max()) <= tol_frac * scale, (t, float(gap.max()), scale)
    return worst, scale


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (the '1e-4 relative' of BASELINE.json is read as relative to the tensor scale)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# --------------------------------------------------------------------------------------------------------------
# numpy restatement of the library's counter-based dropout masks (image2text_b200/csrc/rng.cuh): Philox4x32-10
# (Salmon, Moraes, Dror, Shaw, SC'11; known-answer vectors from Random123's kat_vectors in test_oracle_golden.py)
# and the element -> (counter, word) maps of the three kinds of site.  Test infrastructure only.
# --------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast); returns the four output words."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & mask for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & mask, n2, p0 & mask
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


class DropMasks:
    """Multipliers (0 or 1/(1-p)) of successive dropout sites, for the oracle's `drop=` hooks: the same (seed, offset,
    site) -> mask function the CUDA kernels evaluate.  One call = one site, also when p == 0 (returns None)."""

    def __init__(self, seed: int, offset: int, base: int = 0):
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = ((seed >> 32) ^ (offset >> 32)) & 0xFFFFFFFF
        self.off = offset & 0xFFFFFFFF
        self.n = base

    def _next(self):
        s = self.n
        self.n += 1
        return s

    @staticmethod
    def _thr(p):
        return min(int(p * 4294967296.0 + 0.5), 4294967295)

    def _mult(self, words, p):
        keep = words >= np.uint32(self._thr(p))
        return torch.from_numpy(np.where(keep, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32))

    def elem(self, p, shape):
        site = self._next()
        if p <= 0:
            return None
        n = int(np.prod(shape))
        e4 = np.arange((n + 3) // 4, dtype=np.uint64)
        w = philox4x32_10(e4 & np.uint64(0xFFFFFFFF), e4 >> np.uint64(32), site, self.off, self.k0, self.k1)
        flat = np.stack(w, axis=1).reshape(-1)[:n]
        return self._mult(flat, p).reshape(tuple(shape))

    def tokens(self, p, rows):
        """(rows, 3) multipliers for the q, k, v segments of a packed row (words 0..2 of counter = row)."""
        site = self._next()
        if p <= 0:
            return None
        r = np.arange(rows, dtype=np.uint64)
        w = philox4x32_10(r, 0, site, self.off, self.k0, self.k1)
        return self._mult(np.stack(w[:3], axis=1), p)

    def attn(self, p, B, H, Tq, Tk):
        """csrc/rng.cuh drop_attn8: 16 random bits per key; call (kj >> 4, t >> 1) with t = (kj >> 1) & 3, half-word
        4 (t & 1) + 2 ((kj >> 3) & 1) + (kj & 1); keep iff the half-word >= round(p * 2^16)."""
        site = self._next()
        if p <= 0:
            return None
        row = np.arange(B * H * Tq, dtype=np.uint64)[:, None]
        kj = np.arange(Tk, dtype=np.uint64)[None, :]
        t = (kj >> np.uint64(1)) & np.uint64(3)
        c1 = (kj >> np.uint64(4)) * np.uint64(2) + (t >> np.uint64(1))
        half = ((t & np.uint64(1)) * np.uint64(4) + ((kj >> np.uint64(3)) & np.uint64(1)) * np.uint64(2) + (kj & np.uint64(1))).astype(np.int64)
        w = philox4x32_10(row, c1, site | 0x80000000, self.off, self.k0, self.k1)
        stacked = np.stack(w, axis=-1)                                   # (rows, Tk, 4)
        words = np.take_along_axis(stacked, np.broadcast_to((half >> 1)[..., None], stacked.shape[:2] + (1,)), axis=-1)[..., 0]
        vals = (words.astype(np.uint64) >> (np.uint64(16) * (half & 1).astype(np.uint64))) & np.uint64(0xFFFF)
        thr16 = min(int(p * 65536.0 + 0.5), 65535)
        keep = vals >= np.uint64(thr16)
        mult = torch.from_numpy(np.where(keep, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32))
        return mult.reshape(B, H, Tq, Tk)
