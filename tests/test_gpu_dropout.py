"""Training-mode dropout on the GPU.  The reference draws its masks from torch's generator, so the masks themselves cannot
be compared with it; what is pinned instead:
  * the kernels' masks equal the numpy restatement of the counter-based generator BIT FOR BIT (tests/helpers.DropMasks,
    Philox4x32-10 checked against the Random123 known-answer vectors in the CPU suite), in every attention kernel
    (fp32, mma.sync, tcgen05) and every elementwise site;
  * with those masks handed to the oracle as explicit multipliers (the reference's formulas: nn.Dropout / SDPA dropout_p
    / nn.MultiheadAttention dropout), forward, loss and gradients of kernels and whole models match the oracle;
  * keep rates are Bernoulli(1-p), replays of a captured step draw fresh masks, eval mode is untouched.
All calls go through the C ABI."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from image2text_b200 import VisionEncoderDecoder, ops  # noqa: E402
from image2text_b200._lib import lib  # noqa: E402
from image2text_b200.config_schema import TrainerWrapperConfig  # noqa: E402
from image2text_b200.synthetic import synth_images, synth_labels  # noqa: E402
from image2text_b200.wrapper import ModelTrainerWrapper  # noqa: E402
from oracle import i2t_oracle as O  # noqa: E402
from tests.helpers import SPEC_OVERRIDES, DropMasks, rel_err, spec_and_weights  # noqa: E402

DEV = "cuda"
SEED, OFF = 0x5DEECE66D1234, 7


def state():
    return torch.tensor([SEED, OFF], dtype=torch.int64, device=DEV)


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_elementwise_and_token_masks_equal_numpy_philox(p):
    n = 4 * 12345
    y = rnd(n, seed=1)
    res = rnd(n, seed=2)
    site = ops.DropSite(p, state(), 5)
    want = DropMasks(SEED, OFF, base=5).elem(p, (n,))
    out = ops.dropout_add(y.to(DEV), res.to(DEV), site).cpu()
    assert torch.equal(out, y * want + res) or rel_err(out, y * want + res) < 1e-6
    assert torch.equal(out == res, want == 0)                                   # the mask itself, bit for bit
    out16 = ops.dropout_add(y.to(DEV).bfloat16(), None, site).cpu()
    assert torch.equal(out16 == 0, (want == 0) | (y.bfloat16() == 0))
    g = ops.dropout_bwd(y.to(DEV), torch.bfloat16, site).float().cpu()
    assert rel_err(g, (y * want).bfloat16().float()) < 1e-6
    keep = float((want != 0).float().mean())
    assert abs(keep - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5                   # Bernoulli(1-p)
    # token-level q/k/v masks on a packed (rows, 3*seg) buffer with a row pitch
    rows, seg = 333, 64
    x = torch.ones((rows, 3 * seg + 8), device=DEV)
    ops.token_dropout_(x[:, :3 * seg], seg, 3, ops.DropSite(p, state(), 9))
    tok = DropMasks(SEED, OFF, base=9).tokens(p, rows)
    assert torch.equal(x[:, :3 * seg].cpu(), tok.repeat_interleave(seg, dim=1)) and bool((x[:, 3 * seg:] == 1).all())


def _mask_readout(B, T, H, hs, dtype, site):
    """Attention with q = 0 (uniform probabilities) and one-hot value rows returns the dropout multipliers themselves."""
    C = H * hs
    got = torch.zeros(B, H, T, T)
    for k0 in range(0, T, hs):
        qkv = torch.zeros(B, T, 3 * C)
        for h in range(H):
            for j in range(k0, min(T, k0 + hs)):
                qkv[:, j, 2 * C + h * hs + (j - k0)] = 1.0
        out = ops.attention_packed(qkv.view(B * T, 3 * C).to(DEV).to(dtype), B, T, H, ops.MASK_NONE, 0, drop=site)
        out = out.float().cpu().view(B, T, H, hs).permute(0, 2, 1, 3) * T           # (B,H,T,hs): multiplier of key k0 + e
        w = min(T, k0 + hs) - k0
        got[..., k0:k0 + w] = out[..., :w]
    return got


@pytest.mark.parametrize("kernel", ["fp32", "mma", "tcgen05"])
def test_attention_masks_equal_numpy_philox(kernel):
    B, T, H, hs, p = 2, 150, 2, 64, 0.25
    site = ops.DropSite(p, state(), 3)
    want = DropMasks(SEED, OFF, base=3).attn(p, B, H, T, T)
    lib().i2t_set_tensor_core_attention({"fp32": 0, "mma": 2, "tcgen05": 1}[kernel])
    try:
        got = _mask_readout(B, T, H, hs, torch.float32 if kernel == "fp32" else torch.bfloat16, site)
    finally:
        lib().i2t_set_tensor_core_attention(1)
    assert torch.equal(got != 0, want != 0)
    assert rel_err(got, want) < (1e-5 if kernel == "fp32" else 1e-2)
    assert abs(float((want != 0).float().mean()) - (1 - p)) < 0.01


def _ref_attn(qkv, B, T, H, mode, n_prompt, pmul):
    C = qkv.shape[1] // 3
    q, k, v = qkv.view(B, T, 3 * C).split(C, dim=2)
    i, j = torch.arange(T)[:, None], torch.arange(T)[None, :]
    vis = torch.ones(T, T, dtype=torch.bool) if mode == ops.MASK_NONE else (j <= i)
    if mode == ops.MASK_PROMPT:
        vis = vis & ((i < n_prompt) | (j >= n_prompt))
    mask = torch.zeros(T, T, dtype=qkv.dtype).masked_fill(~vis, -float("inf"))[None, None]
    y = O.sdpa(O.split_heads(q, H), O.split_heads(k, H), O.split_heads(v, H), mask, pmul)
    return O.merge_heads(y).reshape(B * T, C)


@pytest.mark.parametrize("B,T,H,hs,mode,n_prompt", [(2, 256, 4, 64, 2, 8), (2, 197, 3, 64, 0, 0), (3, 70, 4, 32, 1, 0),
                                                    (1, 272, 2, 64, 1, 0)])
def test_attention_dropout_fwd_bwd_match_oracle(B, T, H, hs, mode, n_prompt):
    C, p = H * hs, 0.1
    qkv, dout = rnd(B * T, 3 * C, seed=14), rnd(B * T, C, seed=15)
    site = ops.DropSite(p, state(), 11)
    pmul = DropMasks(SEED, OFF, base=11).attn(p, B, H, T, T).double()
    qr = qkv.double().requires_grad_(True)
    ref = _ref_attn(qr, B, T, H, mode, n_prompt, pmul)
    ref.backward(dout.double())
    out, lse = ops.attention_packed(qkv.to(DEV), B, T, H, mode, n_prompt, want_lse=True, drop=site)
    assert rel_err(out.cpu(), ref) < 3e-6
    dqkv = ops.attention_packed_bwd(qkv.to(DEV), out, dout.to(DEV), lse, B, T, H, mode, n_prompt, drop=site)
    assert rel_err(dqkv.cpu(), qr.grad) < 2e-5
    # bf16: tcgen05 / mma.sync forward, mma.sync backward, against fp64 on the bf16-rounded inputs
    q16, d16 = qkv.to(DEV).bfloat16(), dout.to(DEV).bfloat16()
    qr16 = q16.double().cpu().requires_grad_(True)
    ref16 = _ref_attn(qr16, B, T, H, mode, n_prompt, pmul)
    ref16.backward(d16.double().cpu())
    for tc_mode in (1, 2):
        lib().i2t_set_tensor_core_attention(tc_mode)
        try:
            o16, lse16 = ops.attention_packed(q16, B, T, H, mode, n_prompt, want_lse=True, drop=site)
            g16 = ops.attention_packed_bwd(q16, o16, d16, lse16, B, T, H, mode, n_prompt, drop=site)
        finally:
            lib().i2t_set_tensor_core_attention(1)
        assert rel_err(o16.float().cpu(), ref16) < 2e-2, tc_mode
        assert rel_err(g16.float().cpu(), qr16.grad) < 2e-2, tc_mode


@pytest.mark.parametrize("B,T,S,H,hs", [(2, 256, 8, 12, 64), (3, 20, 4, 4, 32), (1, 33, 16, 12, 64)])
def test_cross_attention_dropout_matches_oracle(B, T, S, H, hs):
    C, p = H * hs, 0.1
    q, kv, dout = rnd(B * T, C, seed=16), rnd(B * S, 2 * C, seed=17), rnd(B * T, C, seed=18)
    site = ops.DropSite(p, state(), 2)
    pmul = DropMasks(SEED, OFF, base=2).attn(p, B, H, T, S).double()

    def ref_fn(qd, kvd):
        k, v = kvd.view(B, S, 2 * C).split(C, dim=2)
        y = O.sdpa(O.split_heads(qd.view(B, T, C), H), O.split_heads(k, H), O.split_heads(v, H), None, pmul)
        return O.merge_heads(y).reshape(B * T, C)
    for dtype, tol_o, tol_g in ((torch.float32, 3e-6, 2e-5), (torch.bfloat16, 2e-2, 2e-2)):
        qd, kvd, dd = q.to(DEV).to(dtype), kv.to(DEV).to(dtype), dout.to(DEV).to(dtype)
        qr, kvr = qd.double().cpu().requires_grad_(True), kvd.double().cpu().requires_grad_(True)
        ref = ref_fn(qr, kvr)
        ref.backward(dd.double().cpu())
        out, lse = ops.xattn_tc(qd, kvd, B, T, S, H, drop=site)
        dq, dkv = ops.xattn_tc_bwd(qd, kvd, out, dd, lse, B, T, S, H, drop=site)
        assert rel_err(out.float().cpu(), ref) < tol_o
        assert rel_err(dq.float().cpu(), qr.grad) < tol_g and rel_err(dkv.float().cpu(), kvr.grad) < tol_g


def _wrapper(name, over, eos, seed=77, **kw):
    tc, _, _ = spec_and_weights(name)
    so = dict(SPEC_OVERRIDES[name], **over)
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    spec = spec_from_config(tc.model, **so)
    sd = synth_state_dict(spec, seed=0)
    tok = types.SimpleNamespace(eos_token_id=eos, bos_token_id=eos, mask_token_id=None, vocab_size=spec["vocab_size"])
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(**kw), -100, device="cuda", spec_overrides=so)
    w.model.load_state_dict(sd)
    w.model.set_dropout_seed(seed)
    return w, spec, sd


def test_tiny_train_step_with_dropout_matches_oracle_with_the_same_masks():
    """Every dropout of the nanoGPT-style decoder on (transformer.drop, token-level q/k/v, SDPA dropout_p, resid_dropout,
    cross-attention dropout, MLP dropout): loss and every gradient against the CPU oracle fed with the same masks."""
    w, spec, sd = _wrapper("tiny", dict(dropout=0.1, attn_dropout=0.2), eos=612)
    w.train()
    images = synth_images(3, 32, seed=11)
    labels = synth_labels(3, 20, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=612)
    sdo = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    tied = "decoder.lm_head.weight"
    sdo[tied] = sdo["decoder.transformer.wte.weight"]
    for step in (1, 2):                                                           # the step offset advances per forward
        loss, _ = w.train_step(images.cuda(), labels.cuda())
        w.model.zero_grad(set_to_none=True)
        loss.backward()
        ids, msk = O.wrapper_inputs(labels, eos_token_id=612, bos_token_id=612)
        for v in sdo.values():
            v.grad = None
        _, logits_o, _ = O.ved_forward(sdo, spec, images, ids, attn_msk=msk, drop=DropMasks(77, step))
        loss_o = O.lm_loss(logits_o, labels, eos_token_id=612)
        loss_o.backward()
        assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o)), step
        checked = 0
        for k, prm in w.model.named_parameters():
            if prm.grad is None or sdo[k].grad is None:
                continue
            assert rel_err(prm.grad.cpu(), sdo[k].grad) < 5e-4, (step, k)
            checked += 1
        assert checked > 40
    # the two steps used different masks; eval mode is the deterministic dropout-free forward
    w.eval()
    with torch.no_grad():
        a = w.model(images=images.cuda(), ids=labels.clamp_min(0).cuda()).logits
        b = w.model(images=images.cuda(), ids=labels.clamp_min(0).cuda()).logits
        _, ref, _ = O.ved_forward(sd, spec, images, labels.clamp_min(0))
    assert torch.equal(a, b) and rel_err(a.cpu(), ref) < 1e-4


def test_bf16_training_forward_with_dropout_tracks_oracle():
    """bf16 compute path (tensor-core attention kernels regenerate the same masks): logits within the bf16 tolerance."""
    tc, _, _ = spec_and_weights("tiny")
    so = dict(SPEC_OVERRIDES["tiny"], dropout=0.1, attn_dropout=0.1)
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    spec = spec_from_config(tc.model, **so)
    sd = synth_state_dict(spec, seed=0)
    m = VisionEncoderDecoder(tc.model, spec_overrides=so, device="cuda", compute_dtype=torch.bfloat16)
    m.load_state_dict(sd)
    m.set_dropout_seed(5)
    m.train()
    images = synth_images(3, 32, seed=11)
    ids = torch.randint(0, 600, (3, 24), generator=torch.Generator().manual_seed(3))
    out = m(images=images.cuda(), ids=ids.cuda())
    _, ref, _ = O.ved_forward(sd, spec, images, ids, drop=DropMasks(5, 1))
    assert rel_err(out.logits.float().cpu(), ref) < 2e-2
    (out.logits.float() ** 2).mean().backward()
    g = dict(m.named_parameters())["decoder.transformer.h.1.mlp.c_fc.weight"].grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0


def test_graphed_micro_steps_draw_fresh_masks_and_match_eager():
    eager, spec, _ = _wrapper("tiny", dict(dropout=0.1, attn_dropout=0.1), eos=612)
    graphed, _, _ = _wrapper("tiny", dict(dropout=0.1, attn_dropout=0.1), eos=612)
    eager.train()
    graphed.train()
    im = synth_images(3, 32, seed=30).cuda()
    lb = synth_labels(3, 20, spec["vocab_size"], seed=40, min_len=3, max_len=14, eos=612).cuda()
    le, lg = [], []
    for _ in range(6):                                     # same batch every time: only the masks change
        l1, _ = eager.train_step(im, lb)
        l1.backward()
        le.append(float(l1))
        lg.append(float(graphed.train_step_graphed(im, lb, 1.0)))
    assert graphed._graph_state["graph"] is not None
    assert len({round(x, 6) for x in lg[2:]}) == 4, lg     # replays are not frozen on the captured masks
    for a, b in zip(le, lg):
        assert abs(a - b) < 1e-5 * abs(a)
    pe, pg = dict(eager.model.named_parameters()), dict(graphed.model.named_parameters())
    for k, prm in pe.items():
        if prm.grad is not None:
            assert rel_err(pg[k].grad.cpu(), prm.grad.cpu()) < 2e-5, k


def test_hf_gpt2_layout_dropout_matches_oracle():
    """HF GPT-2 layout (embd / attn / resid dropouts of GPT2Config, cross attention in every block), small batch."""
    tc, _, _ = spec_and_weights("gpt2")
    so = dict(SPEC_OVERRIDES["gpt2"], dropout=0.1, attn_dropout=0.1)
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    spec = spec_from_config(tc.model, **so)
    sd = synth_state_dict(spec, seed=0)
    m = VisionEncoderDecoder(tc.model, device="cuda", spec_overrides=so)
    m.load_state_dict(sd)
    m.set_dropout_seed(9)
    m.train()
    images = synth_images(1, 224, seed=31)
    ids = torch.randint(0, 50000, (1, 40), generator=torch.Generator().manual_seed(4))
    out = m(images=images.cuda(), ids=ids.cuda())
    keys = ["decoder.backbone.transformer.h.0.attn.c_attn.weight", "decoder.backbone.transformer.h.3.crossattention.q_attn.weight",
            "decoder.backbone.transformer.h.11.mlp.c_proj.bias", "decoder.backbone.transformer.wpe.weight"]
    probe = torch.randn(out.logits.shape, generator=torch.Generator().manual_seed(5))
    (out.logits.float() * probe.cuda()).sum().backward()
    sdo = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    _, logits_o, _ = O.ved_forward(sdo, spec, images, ids, drop=DropMasks(9, 1))
    assert rel_err(out.logits.cpu(), logits_o) < 2e-4
    (logits_o * probe).sum().backward()
    named = dict(m.named_parameters())
    for k in keys:
        assert rel_err(named[k].grad.cpu(), sdo[k].grad) < 1e-3, k


@pytest.mark.parametrize("B,T,H,dtype", [(2, 256, 3, torch.bfloat16), (2, 197, 2, torch.bfloat16), (1, 272, 2, torch.bfloat16),
                                         (2, 70, 2, torch.float32)])
def test_token_dropout_backward_rides_on_the_attention_backward(B, T, H, dtype):
    """attention_packed_bwd(tok=site) == attention_packed_bwd followed by token_dropout_ on the packed gradient: the tcgen05 kernel
    scales dq / dk / dv by the (row, segment) masks while it stores them (<= 256 rows), every other kernel is followed by the
    stand-alone pass inside the same C call."""
    C = H * 64
    qkv = rnd(B * T, 3 * C, seed=21).to(DEV).to(dtype)
    dout = rnd(B * T, C, seed=22).to(DEV).to(dtype)
    st = state()
    drop, tok = ops.DropSite(0.1, st, 5), ops.DropSite(0.2, st, 6)
    out, lse = ops.attention_packed(qkv, B, T, H, ops.MASK_CAUSAL, 0, want_lse=True, drop=drop)
    want = ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, ops.MASK_CAUSAL, 0, drop=drop)
    ops.token_dropout_(want, C, 3, tok)
    got = ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, ops.MASK_CAUSAL, 0, drop=drop, tok=tok)
    zeros = (want == 0).all(dim=0).sum()                    # whole columns are never zero; whole (row, segment) blocks are
    assert float((want.float().abs().sum(dim=1) == 0).float().mean()) < 0.05 and int(zeros) == 0
    assert rel_err(got.float().cpu(), want.float().cpu()) < (1e-2 if dtype == torch.bfloat16 else 1e-6)
    assert torch.equal(got == 0, want == 0)                 # exactly the same (row, segment) blocks are dropped
