"""Decode megakernel v3 (bf16, one launch per generate call, dataflow instead of grid barriers): parity against the ORACLE.

A bf16 decode accumulates in another order than any other bf16 execution, so ids are not compared token by token with
one (a near-tie may legitimately flip).  Every token the megakernel picked is checked against the fp32 CPU oracle
(oracle/i2t_oracle.py, pinned to the unmodified reference by tests/golden) run TEACHER-FORCED over the generated prefix:
  greedy : the pick is not banned by the no-repeat-n-gram rule and its oracle logit is within 2e-2 * max|logit|
           (BASELINE.json's bf16 tolerance) of the oracle's best non-banned logit;
  top-k  : the pick is not banned and its oracle logit is within the tolerance of the oracle's k-th best non-banned one.
The oracle is given the encoder output the bf16 model computed (the LSH tail is an integer hash of the ViT feature: a
bf16 feature may land in a neighbouring bucket; the bf16 ViT trunk itself is pinned in test_gpu_model.py).
The fp32 bit-exact parity of the decode arithmetic is tests/test_gpu_model.py (kernels / mega modes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402
from tests.helpers import check_picks_vs_oracle, spec_and_weights  # noqa: E402,F401
from tests.test_gpu_model import build  # noqa: E402


def test_mega3_is_the_default_bf16_engine():
    m = build("nano", torch.bfloat16)
    eng = DecodeEngine(m, 8)
    assert eng.mode == "mega3"
    assert DecodeEngine(build("nano"), 8).mode == "kernels"        # fp32: the parity anchor keeps the separate kernels


def test_mega3_tiny_prefill_small_batch_and_topk():
    m = build("tiny", torch.bfloat16)
    spec = m.spec
    images = synth_images(3, 32, seed=11).cuda()
    eos = spec["vocab_size"] - 1
    g = torch.Generator().manual_seed(5)
    prompt = torch.cat([torch.full((3, 1), eos), torch.randint(0, eos, (3, 3), generator=g)], dim=1).cuda()
    eng = DecodeEngine(m, 3, mode="mega3")
    assert eng.mode == "mega3"
    got = eng.generate(images, prompt, 24, 1.0, 1, seed=0)           # 3 prefill steps + 24 sampled steps, one launch
    assert eng.launches_per_step > 2                                 # first call: one pack launch per linear op + the loop
    assert torch.equal(got[:, :4], prompt)
    assert int(got.min()) >= 0 and int(got.max()) < spec["vocab_size"]
    assert torch.equal(got, eng.generate(images, prompt, 24, 1.0, 1, seed=0))          # deterministic (fixed reduction order)
    assert eng.launches_per_step == 2                                # one poison-fill launch + ONE launch for the whole token loop
    check_picks_vs_oracle("tiny", m, images, got, 4, top_k=1)
    sampled = eng.generate(images, prompt, 20, 0.8, 5, seed=7)
    assert torch.equal(sampled, eng.generate(images, prompt, 20, 0.8, 5, seed=7))      # same seed -> same draw
    assert not torch.equal(sampled, eng.generate(images, prompt, 20, 0.8, 5, seed=8))
    check_picks_vs_oracle("tiny", m, images, sampled, 4, top_k=5)


def test_mega3_tiny_agrees_with_mega2():
    """Same arithmetic (mma tiles, K split over 8 warps, fixed-order reduction), different plumbing: ids agree unless a
    near-tie flips on the one-pass LayerNorm statistics -- report, require the first tokens."""
    m = build("tiny", torch.bfloat16)
    images = synth_images(3, 32, seed=11).cuda()
    eos = m.spec["vocab_size"] - 1
    prompt = torch.full((3, 1), eos, dtype=torch.long, device="cuda")
    a = DecodeEngine(m, 3, mode="mega3").generate(images, prompt, 24, 1.0, 1, seed=0)
    b = DecodeEngine(m, 3, mode="mega2").generate(images, prompt, 24, 1.0, 1, seed=0)
    same = (a == b).all(dim=1)
    print("rows identical to mega2 over 24 tokens:", int(same.sum()), "of 3")
    assert torch.equal(a[:, :4], b[:, :4])


def test_mega3_nano_bench_workload_teacher_forced(golden):
    """BASELINE.json configs[1] as bench.py runs it: 8 captions x 64 new tokens, greedy, bf16."""
    m = build("nano", torch.bfloat16)
    images = synth_images(8, 224, seed=1234).cuda()
    prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
    eng = DecodeEngine(m, 8, mode="mega3")
    got = eng.generate(images, prompt, 64, 1.0, 1, seed=0)
    assert got.shape == (8, 65) and int(eng.pos.item()) == 64
    again = eng.generate(images, prompt, 64, 1.0, 1, seed=0)
    assert eng.launches_per_step == 2                                # one poison-fill launch + ONE launch for the whole token loop
    assert torch.equal(got, again)
    worst, scale = check_picks_vs_oracle("nano", m, images, got, 1, top_k=1)
    # how far the bf16 ids follow the reference's fp32 greedy ids (informational: a near-tie ends the common prefix)
    ref = torch.from_numpy(np.asarray(golden("nano_generate")["greedy"]))
    agree = (got.cpu() == ref).long().cumprod(dim=1).sum(dim=1) - 1
    print(f"bf16 mega3 vs reference fp32 greedy ids: common prefix per row {agree.tolist()} of 64; worst oracle gap "
          f"{worst:.4f} at logit scale {scale:.2f}")


def test_mega3_nano_topk16_teacher_forced():
    """The notebook's sampling setting (top_k = 16, T = 1.0): every draw inside the oracle's top-16 (2e-2 tolerance)."""
    m = build("nano", torch.bfloat16)
    images = synth_images(8, 224, seed=1234).cuda()
    prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
    eng = DecodeEngine(m, 8, mode="mega3")
    got = eng.generate(images, prompt, 24, 1.0, 16, seed=5)
    assert torch.equal(got, eng.generate(images, prompt, 24, 1.0, 16, seed=5))
    check_picks_vs_oracle("nano", m, images, got, 1, top_k=16)


def test_mega3_four_way_split_with_combine_stage(monkeypatch):
    """I2T_M3_FOLD=0: the down projection as four K = 768 partial ops + a combine stage (the default folds the first K half into
    the epilogue of the second).  Same oracle bar; both variants are deterministic."""
    monkeypatch.setenv("I2T_M3_FOLD", "0")
    m = build("nano", torch.bfloat16)
    images = synth_images(8, 224, seed=1234).cuda()
    prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
    eng = DecodeEngine(m, 8, mode="mega3")
    got = eng.generate(images, prompt, 20, 1.0, 1, seed=0)
    assert torch.equal(got, eng.generate(images, prompt, 20, 1.0, 1, seed=0))
    check_picks_vs_oracle("nano", m, images, got, 1, top_k=1)


def test_mega3_repacks_after_a_weight_update():
    """The packed weight streams follow the fp32 masters: an in-place change of a decoder weight changes the ids."""
    m = build("tiny", torch.bfloat16)
    images = synth_images(3, 32, seed=11).cuda()
    eos = m.spec["vocab_size"] - 1
    prompt = torch.full((3, 1), eos, dtype=torch.long, device="cuda")
    eng = DecodeEngine(m, 3, mode="mega3")
    before = eng.generate(images, prompt, 16, 1.0, 1, seed=0)
    w = m.weights()["decoder.transformer.h.0.mlp.c_fc.weight"]
    saved = w.detach().clone()
    try:
        with torch.no_grad():
            w.mul_(-1.0)
        after = eng.generate(images, prompt, 16, 1.0, 1, seed=0)
        assert not torch.equal(before, after)
    finally:
        with torch.no_grad():
            w.copy_(saved)
    assert torch.equal(before, eng.generate(images, prompt, 16, 1.0, 1, seed=0))


def test_mega3_tcgen05_linear_stages_match_oracle(monkeypatch):
    """The experimental tcgen05 instantiation (I2T_M3_TC=1): layer projections as tcgen05.mma M128 N16 K16 over weights packed as
    128B-swizzled UMMA atoms, accumulators in TMEM, four issuing warps -- same oracle parity as the default mma.sync tiles."""
    monkeypatch.setenv("I2T_M3_TC", "1")
    for name, B, size, P in (("tiny", 3, 32, 2), ("nano", 8, 224, 1)):
        m = build(name, torch.bfloat16)
        images = synth_images(B, size, seed=11).cuda()
        eos = m.spec["vocab_size"] - 1
        g = torch.Generator().manual_seed(5)
        prompt = torch.cat([torch.full((B, 1), eos), torch.randint(0, eos, (B, P - 1), generator=g)], dim=1).cuda()
        eng = DecodeEngine(m, B, mode="mega3")
        got = eng.generate(images, prompt, 20, 1.0, 1, seed=0)
        assert eng._mega3["tc"] == 1
        assert torch.equal(got, eng.generate(images, prompt, 20, 1.0, 1, seed=0))
        check_picks_vs_oracle(name, m, images, got, P, top_k=1)
