"""Training-path parity on the GPU: the B200 ModelTrainerWrapper (forward + fused loss + backward kernels) against
gradients produced by the UNMODIFIED reference (tests/golden/tiny_train.npz, nano_train.npz), the fused optimisers
against torch.optim.AdamW / the oracle's SNRAdam on the same gradients, and the fused EMA against the reference rule."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image2text_b200.config_schema import TrainerWrapperConfig  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.optimizer import AdamW, SNRAdam  # noqa: E402
from image2text_b200.synthetic import synth_images, synth_labels  # noqa: E402
from image2text_b200.wrapper import ModelTrainerWrapper  # noqa: E402
from oracle import i2t_oracle as O  # noqa: E402
from tests.helpers import SPEC_OVERRIDES, rel_err, spec_and_weights  # noqa: E402


def T(x):
    return torch.from_numpy(np.asarray(x))


def make_wrapper(name, trainer_kw, eos):
    tc, spec, sd = spec_and_weights(name)
    tok = types.SimpleNamespace(eos_token_id=eos, bos_token_id=eos, mask_token_id=None, vocab_size=spec["vocab_size"])
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(**trainer_kw), -100, device="cuda",
                            spec_overrides=SPEC_OVERRIDES[name])
    w.model.load_state_dict(sd)
    return w, spec, sd


@pytest.mark.parametrize("name", ["plain", "moco"])
def test_tiny_train_step_matches_reference(golden, name):
    g = golden("tiny_train")
    kw = {} if name == "plain" else dict(moco_momentum=0.9, moco_alpha=0.4, weight_fn="inverse_sqrt_position",
                                          eos_token_weight=2.0, training_temperature=1.3)
    w, spec, sd = make_wrapper("tiny", kw, eos=612)
    if name == "moco":
        w.model_m.load_state_dict(synth_state_dict(spec, seed=7))
    w.train()
    images = synth_images(3, 32, seed=11).cuda()
    labels = synth_labels(3, 20, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=612).cuda()
    loss, metrics = w.train_step(images, labels)
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 1e-4 * abs(float(g[f"{name}_loss"]))
    assert f"train_loss_lm" in metrics
    loss.backward()
    named = dict(w.model.named_parameters())
    checked = 0
    for key, val in g.items():
        if key.startswith(f"{name}_gnorm::"):
            k = key.split("::")[1]
            gr = named[k].grad
            gn = 0.0 if gr is None else float(gr.norm())
            assert abs(gn - float(val)) <= 2e-4 * max(float(val), 1e-7), (k, gn, float(val))
            checked += 1
        if key.startswith(f"{name}_grad::"):
            k = key.split("::")[1]
            assert rel_err(named[k].grad.cpu(), T(val)) < 2e-4, k
    assert checked > 50
    if name == "moco":
        pm = dict(w.model_m.named_parameters())
        for k in ("decoder.transformer.h.0.attn.c_attn.weight", "encoder.model.encoder.ln.bias"):
            assert rel_err(pm[k].cpu(), T(g[f"moco_ema::{k}"])) < 1e-6, k
    with torch.no_grad():
        w.eval()
        vloss, vm = w.val_step(images, labels)
    assert abs(float(vloss) - float(g[f"{name}_val_loss"])) < 1e-4 * abs(float(vloss)) and "val_loss_lm" in vm


def test_nano_train_loss_and_grads_match_reference(golden):
    g = golden("nano_train")
    w, spec, sd = make_wrapper("nano", {}, eos=50256)
    w.train()
    images = synth_images(2, 224, seed=21).cuda()
    labels = T(g["labels"]).cuda()
    loss, _ = w.train_step(images, labels)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    named = dict(w.model.named_parameters())
    for key, val in g.items():
        if key.startswith("grad::"):
            k = key.split("::")[1]
            assert rel_err(named[k].grad.cpu(), T(val)) < 5e-4, k
        if key.startswith("gnorm::"):
            k = key.split("::")[1]
            gn = float(named[k].grad.norm())
            assert abs(gn - float(val)) <= 1e-3 * max(float(val), 1e-7), (k, gn, float(val))
    # the frozen ViT trunk (refine_base_model False / LSH tail) gets no gradient, like the reference
    assert named["encoder.model.conv_proj.weight"].grad is None


@pytest.mark.parametrize("cls,ref", [(AdamW, "adamw"), (SNRAdam, "snradam")])
def test_fused_optimizer_classes(cls, ref):
    gen = torch.Generator().manual_seed(3)
    shapes = [(257, 33), (768,), (5,), (64, 64)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=gen).cuda()) for s in shapes]
    cpu = [p.detach().cpu().clone() for p in ps]
    groups = [dict(params=ps[:2], lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1), dict(params=ps[2:], lr=1e-3, betas=(0.9, 0.999))]
    opt = cls(groups)
    hp = [(3e-3, 0.9, 0.95, 0.1)] * 2 + [(1e-3, 0.9, 0.999, 0.0)] * 2
    ms = [torch.zeros_like(c) for c in cpu]
    vs = [torch.zeros_like(c) for c in cpu]
    for step in range(1, 4):
        grads = [torch.randn(s, generator=gen) * step for s in shapes]
        for p, gr in zip(ps, grads):
            p.grad = gr.cuda()
        opt.step()
        opt.zero_grad(set_to_none=False)
        for i in range(4):
            lr, b1, b2, wd = hp[i]
            fn = O.adamw_step if ref == "adamw" else O.snradam_step
            fn(cpu[i], grads[i], ms[i], vs[i], step, lr, b1, b2, 1e-8, wd)
            assert rel_err(ps[i].detach().cpu(), cpu[i]) < 3e-6, (step, i)
    sd = opt.state_dict()
    assert len(sd["state"]) == 4 and "exp_avg" in sd["state"][0]


def test_gpt2_hf_layout_forward_and_backward_match(golden):
    """BASELINE configs[2]/[3] model: ViT-B/16 (trainable) + per-slot MLP tail + HF-layout GPT-2 with cross attention.
    Forward against the reference-made fixture, gradients against the CPU oracle's autograd on the same weights."""
    from image2text_b200 import VisionEncoderDecoder
    g = golden("gpt2_fwd")
    tc, spec, sd = spec_and_weights("gpt2")
    m = VisionEncoderDecoder(tc.model, device="cuda", spec_overrides=SPEC_OVERRIDES["gpt2"])
    m.load_state_dict(sd)
    images = synth_images(2, 224, seed=31)
    labels = T(g["labels"])
    ids = torch.where(labels != -100, labels, torch.full_like(labels, 50256))
    m.eval()
    with torch.no_grad():
        out = m(images=images.cuda(), ids=ids.cuda())
    assert rel_err(out.encoder_output.cpu(), T(g["enc"])) < 1e-4
    assert rel_err(out.hidden_state.cpu(), T(g["hidden"])) < 1e-4
    assert rel_err(out.logits.cpu()[..., :256], T(g["logits_head"])) < 1e-4
    assert np.array_equal(out.logits.argmax(-1).cpu().numpy(), g["logits_argmax"])
    # gradients of a scalar loss w.r.t. a few tensors of every kind (Conv1D, cross attention, tail MLP, ViT trunk)
    keys = ["decoder.backbone.transformer.h.0.attn.c_attn.weight", "decoder.backbone.transformer.h.3.crossattention.q_attn.weight",
            "decoder.backbone.transformer.h.11.mlp.c_proj.bias", "decoder.backbone.transformer.h.5.crossattention.c_attn.weight",
            "encoder.proj.models.3.model.0.weight", "encoder.model.encoder.layers.encoder_layer_11.mlp.3.weight",
            "decoder.backbone.transformer.wpe.weight", "encoder.model.conv_proj.bias"]
    m.train()
    out = m(images=images.cuda(), ids=ids.cuda())
    probe = torch.randn(out.logits.shape, generator=torch.Generator().manual_seed(5)).cuda()
    (out.logits.float() * probe).sum().backward()
    named = dict(m.named_parameters())
    sdo = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    _, logits_o, _ = O.ved_forward(sdo, spec, images, ids)
    (logits_o * probe.cpu()).sum().backward()
    for k in keys:
        assert rel_err(named[k].grad.cpu(), sdo[k].grad) < 5e-4, k


def test_graphed_micro_step_accumulates_the_same_gradients():
    """train_step_graphed (forward + backward of one micro-step replayed as a CUDA graph, loss scale folded into the loss
    kernel's dlogits) must leave the same accumulated .grad as `train_step` + `(loss * scale).backward()`."""
    eager, spec, sd = make_wrapper("tiny", {}, eos=612)
    graphed, _, _ = make_wrapper("tiny", {}, eos=612)
    eager.train()
    graphed.train()
    scale = 0.25
    batches = [(synth_images(3, 32, seed=30 + i).cuda(),
                synth_labels(3, 20, spec["vocab_size"], seed=40 + i, min_len=3, max_len=14, eos=612).cuda()) for i in range(5)]
    losses_e, losses_g = [], []
    for i, (im, lb) in enumerate(batches):
        le, _ = eager.train_step(im, lb)
        (le * scale).backward()
        losses_e.append(float(le))
        lg = graphed.train_step_graphed(im, lb, scale)       # calls 1-2 eager, call 3 captures, 3-5 replay
        losses_g.append(float(lg))
    assert graphed._graph_state["graph"] is not None
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) < 1e-5 * abs(a)
    pe, pg = dict(eager.model.named_parameters()), dict(graphed.model.named_parameters())
    checked = 0
    for k, p in pe.items():
        if p.grad is None:
            continue
        assert rel_err(pg[k].grad.cpu(), p.grad.cpu()) < 2e-5, k
        checked += 1
    assert checked > 20


def test_grad_sinks_add_the_same_gradients_in_place():
    """ops.grad_sinks(): the backward kernels add weight / bias / LayerNorm / wpe / LSH-table gradients straight into existing
    .grad buffers (no AccumulateGrad add, no slice_backward of the packed in_proj weight).  Same accumulated result as plain
    autograd over two micro-steps, and every parameter is announced exactly once per backward."""
    from image2text_b200 import ops
    plain, spec, sd = make_wrapper("tiny", {}, eos=612)
    sunk, _, _ = make_wrapper("tiny", {}, eos=612)
    plain.train()
    sunk.train()
    batches = [(synth_images(3, 32, seed=50 + i).cuda(),
                synth_labels(3, 20, spec["vocab_size"], seed=60 + i, min_len=3, max_len=14, eos=612).cuda()) for i in range(3)]
    for i, (im, lb) in enumerate(batches):
        lp, _ = plain.train_step(im, lb)
        lp.backward()
        if i == 0:                                   # the first backward creates the .grad buffers
            ls, _ = sunk.train_step(im, lb)
            ls.backward()
            continue
        told = []
        with ops.grad_sinks(notify=told.append):
            ls, _ = sunk.train_step(im, lb)
            ls.backward()
        assert float(ls) == float(lp)
        assert len(told) == len({id(p) for p in told}) and len(told) > 20
    pp, ps = dict(plain.model.named_parameters()), dict(sunk.model.named_parameters())
    sunk_ids = {id(p) for p in told}
    n = 0
    for k, p in pp.items():
        if p.grad is None:
            continue
        assert rel_err(ps[k].grad.cpu(), p.grad.cpu()) < 2e-5, k
        n += id(ps[k]) in sunk_ids
    assert n > 20


@pytest.mark.parametrize("graphed", [False, True])
def test_bf16_training_sees_the_updated_weights(graphed):
    """The fused optimiser / EMA kernels write fp32 masters through raw pointers; the bf16 copies every bf16 GEMM and the decode
    engines read must follow (refreshed IN PLACE: CUDA graphs and decode tables hold their addresses).  On ONE fixed batch the
    loss must fall over optimiser steps, every bf16 copy must equal its master rounded to bf16, and the teacher's copies must
    follow the EMA (they used to stay at the initial weights)."""
    tc, spec, sd = spec_and_weights("tiny")
    tok = types.SimpleNamespace(eos_token_id=612, bos_token_id=612, mask_token_id=None, vocab_size=spec["vocab_size"])
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(moco_momentum=0.9, moco_alpha=0.4), -100, device="cuda",
                            compute_dtype=torch.bfloat16, spec_overrides=SPEC_OVERRIDES["tiny"])
    w.model.load_state_dict(sd)
    w.copy_momentum_params()
    w.train()
    opt = AdamW([dict(params=[p for p in w.model.parameters() if p.requires_grad], lr=3e-3, betas=(0.9, 0.95), weight_decay=0.0)])
    images = synth_images(4, 32, seed=11).cuda()
    labels = synth_labels(4, 20, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=612).cuda()
    addr0 = {k: v._i2t_shadow.data_ptr() for k, v in w.model._tensors().items() if getattr(v, "_i2t_shadow", None) is not None}
    losses = []
    for step in range(12):
        if graphed:
            loss = w.train_step_graphed(images, labels, 1.0)
        else:
            loss, _ = w.train_step(images, labels)
            loss.backward()
        losses.append(float(loss))
        opt.step()
        opt.zero_grad(set_to_none=False)
        if step == 0:
            addr0 = {k: v._i2t_shadow.data_ptr() for k, v in w.model._tensors().items() if getattr(v, "_i2t_shadow", None) is not None}
    assert losses[-1] < losses[0] - 0.5, losses
    for mdl in (w.model, w.model_m):
        n = 0
        for k, v in mdl._tensors().items():
            sh = getattr(v, "_i2t_shadow", None)
            if sh is not None:
                assert torch.equal(sh, v.detach().to(torch.bfloat16)), k
                n += 1
        assert n > 20
    for k, a in addr0.items():                    # refreshed in place: same addresses as after the first step
        assert w.model._tensors()[k]._i2t_shadow.data_ptr() == a, k
    # the teacher moved away from its initial copy of the student (EMA) and its bf16 copies moved with it
    k = "decoder.transformer.h.0.mlp.c_fc.weight"
    assert not torch.equal(w.model_m._tensors()[k].detach().cpu(), sd[k])
    # generate() after training uses the trained weights (tables are rebuilt / copies refreshed)
    w.eval()
    prompt = torch.full((4, 1), 612, dtype=torch.long, device="cuda")
    a = w.model.generate(images, prompt, max_new_tokens=8, top_k=1)
    fresh = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(), -100, device="cuda", compute_dtype=torch.bfloat16,
                                spec_overrides=SPEC_OVERRIDES["tiny"]).model
    fresh.load_state_dict(w.model.state_dict())
    fresh.eval()
    assert torch.equal(a, fresh.generate(images, prompt, max_new_tokens=8, top_k=1))


def test_contrastive_auxiliary_loss_matches_reference(golden):
    """training/wrapper.py:98-118,206-209 (add_contrastive_loss): LM + contrastive loss and every gradient norm against the
    reference-made fixture tests/golden/tiny_contrastive.npz (similarity GEMM + the masked, weighted CE kernel)."""
    g = golden("tiny_contrastive")
    kw = dict(add_contrastive_loss=True, training_contrastive_temperature=0.7, weight_fn="inverse_sqrt_position", eos_token_weight=2.0)
    w, spec, sd = make_wrapper("tiny", kw, eos=612)
    w.train()
    images = synth_images(3, 32, seed=11).cuda()
    labels = T(g["labels"]).cuda()
    loss, metrics = w.train_step(images, labels)
    assert abs(float(metrics["train_loss_lm"]) - float(g["loss_lm"])) < 1e-4 * abs(float(g["loss_lm"]))
    assert abs(float(metrics["train_loss_contrastive"]) - float(g["loss_contrastive"])) < 1e-4 * abs(float(g["loss_contrastive"]))
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    named = dict(w.model.named_parameters())
    n = 0
    for key, val in g.items():
        if key.startswith("gnorm::"):
            k = key.split("::")[1]
            if k not in named:
                continue
            gn = float(named[k].grad.norm())
            assert abs(gn - float(val)) <= 3e-4 * max(float(val), 1e-7), (k, gn, float(val))
            n += 1
        elif key.startswith("grad::"):
            assert rel_err(named[key.split("::")[1]].grad.cpu(), T(val)) < 3e-4, key
    assert n > 50
