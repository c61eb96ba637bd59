"""Batched (more than 8 sequences) and HF GPT-2 layout decode in bf16: the step is a programmatic-dependent-launch chain
(image2text_b200/decode_engine.py `_gemm_step` / `_layers_chain`).  Teacher-forced parity as in test_gpu_decode_mega2.py:
every pick must be an un-banned (near-)arg-max / top-k member of the model's own full forward over the generated prefix."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402
from tests.test_gpu_decode_mega2 import check_picks  # noqa: E402
from tests.test_gpu_model import build  # noqa: E402


def test_batched_pdl_chain_teacher_forced():
    """More than 8 sequences in bf16: the decode step is the programmatic-dependent-launch chain (split-K tensor-core
    projections, LayerNorm that zero-fills their output, attention with the fused K/V append, one-pass greedy arg-max).
    Same teacher-forced criterion; a prompt of 3 tokens exercises the prefill steps; a second call reuses the graph."""
    m = build("nano", torch.bfloat16)
    B = 24
    images = synth_images(3, 224, seed=77).cuda().repeat_interleave(8, dim=0)
    g = torch.Generator().manual_seed(9)
    prompt = torch.cat([torch.full((B, 1), 50256), torch.randint(0, 50256, (B, 2), generator=g)], dim=1).cuda()
    eng = DecodeEngine(m, B)
    assert eng.mode == "gemm"
    got = eng.generate(images, prompt, 40, 1.0, 1, seed=0)
    assert got.shape == (B, 43) and torch.equal(got[:, :3], prompt)
    check_picks(m, images, got, 3, top_k=1)
    again = eng.generate(images, prompt, 40, 1.0, 1, seed=0)          # graph replay of the same step
    check_picks(m, images, again, 3, top_k=1)
    # top-k sampling: the split-K projections add their partial tiles with fp32 atomics (summation order not fixed), so two
    # runs with the same seed agree only up to near-ties of the inverse-CDF draw -- both must satisfy the top-k criterion
    for _ in range(2):
        sampled = eng.generate(images, prompt, 24, 0.8, 5, seed=3)
        assert torch.equal(sampled[:, :3], prompt)
        check_picks(m, images, sampled, 3, top_k=5)


def test_hf_gpt2_pdl_chain_teacher_forced():
    """HF GPT-2 layout decoder in bf16 (Conv1D weights as MN-major split-K operands, soft-prompt rows pushed through the
    cache by the same chain): every greedy pick is an un-banned (near-)arg-max of the model's own full forward."""
    m = build("gpt2", torch.bfloat16)
    images = synth_images(4, 224, seed=33).cuda()
    prompt = torch.tensor([[50256, 11, 257]] * 4, dtype=torch.long, device="cuda")
    got = m.generate(images, prompt, max_new_tokens=20, temperature=1.0, top_k=1)
    assert got.shape == (4, 23) and torch.equal(got[:, :3], prompt)
    check_picks(m, images, got, 3, top_k=1)
    check_picks(m, images, m.generate(images, prompt, max_new_tokens=20, temperature=1.0, top_k=1), 3, top_k=1)   # graph replays
