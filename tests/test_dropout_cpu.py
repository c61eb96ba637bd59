"""CPU checks of the dropout-mask restatement (tests/helpers.py) that the GPU tests use as their checker: Philox4x32-10
against the published known-answer vectors of Random123 (kat_vectors: philox4x32 10), and the basic properties of the
three element -> (counter, word) maps."""
import numpy as np
import torch

from oracle import i2t_oracle as O
from tests.helpers import DropMasks, philox4x32_10

KATS = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_known_answers():
    for ctr, key, want in KATS:
        got = [int(x) for x in philox4x32_10(*[np.uint32(c) for c in ctr], *key)]
        assert got == list(want)
    # vectorised evaluation equals element-wise evaluation
    c0 = np.arange(7, dtype=np.uint32)
    vec = philox4x32_10(c0, 5, 9, 1, 123, 456)
    for i in range(7):
        one = philox4x32_10(np.uint32(i), 5, 9, 1, 123, 456)
        assert [int(w[i]) for w in vec] == [int(w) for w in one]


def test_mask_maps():
    p = 0.3
    a = DropMasks(99, 4).elem(p, (5, 7, 12))
    b = DropMasks(99, 4).elem(p, (5 * 7 * 12,))
    assert torch.equal(a.reshape(-1), b) and set(a.unique().tolist()) <= {0.0, np.float32(1 / (1 - p)).item()}
    assert not torch.equal(a, DropMasks(99, 5).elem(p, (5, 7, 12)))            # the step offset changes the mask
    assert not torch.equal(a, DropMasks(99, 4, base=1).elem(p, (5, 7, 12)))    # so does the site
    d = DropMasks(1, 1)
    assert d.elem(0.0, (4,)) is None and d.n == 1                               # p = 0 still consumes a site index
    t = DropMasks(3, 0).tokens(0.5, 1000)
    assert t.shape == (1000, 3) and abs(float((t != 0).float().mean()) - 0.5) < 0.05
    m = DropMasks(3, 0).attn(0.2, 2, 3, 17, 40)
    assert m.shape == (2, 3, 17, 40) and abs(float((m != 0).float().mean()) - 0.8) < 0.03
    # a shorter key range is a prefix of a longer one (the mask of (row, key) does not depend on Tk)
    m2 = DropMasks(3, 0).attn(0.2, 1, 1, 17, 64)
    m3 = DropMasks(3, 0).attn(0.2, 1, 1, 17, 33)
    assert torch.equal(m2[..., :33], m3)


def test_oracle_dropout_hooks_are_identity_for_p_zero_and_scale_by_the_multiplier():
    q, k, v = (torch.randn(2, 3, 5, 8, generator=torch.Generator().manual_seed(i)) for i in range(3))
    base = O.sdpa(q, k, v, None)
    assert torch.equal(O.sdpa(q, k, v, None, None), base)
    assert torch.allclose(O.sdpa(q, k, v, None, torch.full((2, 3, 5, 5), 2.0)), 2.0 * base, atol=1e-6)
    drop_first_key = torch.ones(2, 3, 5, 5)
    drop_first_key[..., 0] = 0
    s = (q @ k.transpose(-1, -2)) / 8 ** 0.5
    pr = torch.softmax(s, -1)
    pr[..., 0] = 0
    assert torch.allclose(O.sdpa(q, k, v, None, drop_first_key), pr @ v, atol=1e-6)
