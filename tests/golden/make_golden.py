"""Generate the golden fixtures by running the UNMODIFIED reference on CPU.

    python tests/golden/make_golden.py            # needs /root/reference (read-only) in this container

Outputs ``tests/golden/*.npz`` (committed).  Weights are never stored: both sides rebuild them with
``image2text_b200.model_spec.synth_state_dict(spec, seed)``; the reference loads them through its own
``load_state_dict`` (strict), which also pins the checkpoint key layout.
"""
import os
import sys
import time

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from image2text_b200.config_schema import load_training_config  # noqa: E402
from image2text_b200.model_spec import spec_from_config, state_schema, synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images, synth_labels  # noqa: E402


def small_vit_patch(ref, layers, image):
    """Shrink the torchvision trunk the reference instantiates (tiny parity config only)."""
    from torchvision.models.vision_transformer import VisionTransformer

    def _vit(weights=None):
        return VisionTransformer(image_size=image, patch_size=16, num_layers=layers, num_heads=12,
                                 hidden_dim=768, mlp_dim=3072)
    _vit._i2t_offline = True
    ref.E.vit_b_16 = _vit


def restore_vit_patch(ref):
    import torchvision

    def _vit(weights=None):
        return torchvision.models.vit_b_16(weights=None)
    _vit._i2t_offline = True
    ref.E.vit_b_16 = _vit


def build(ref, cfg_name, spec_over, zero_dropout=True):
    tc = load_training_config(os.path.join(ROOT, "configs", cfg_name + ".yaml"))
    spec = spec_from_config(tc.model, **spec_over)
    raw = yaml.safe_load(open(os.path.join(ROOT, "configs", cfg_name + ".yaml")))
    mdl = raw["model"]
    if zero_dropout and "transformer_config" in mdl["decoder_config"]:
        mdl["decoder_config"]["transformer_config"]["attn_config"]["dropout"] = 0.0
        mdl["decoder_config"]["transformer_config"]["attn_config"]["attn_dropout"] = 0.0
    if "pretrained_model" in mdl["decoder_config"]:
        mdl["decoder_config"]["pretrained_model"] = None     # random-init TransformerDecoder, no hub download
    mdl["decoder_config"].pop("lora_spec", None)
    sd = synth_state_dict(spec, seed=0)
    model = ref_harness.build_reference_model(mdl, state_dict=None)
    ref_sd = model.state_dict()
    schema = state_schema(spec)
    assert list(sorted(ref_sd.keys())) == list(sorted(schema.keys())), \
        (sorted(set(ref_sd) ^ set(schema)))
    for k, (shape, dtype) in schema.items():
        assert tuple(ref_sd[k].shape) == tuple(shape) and ref_sd[k].dtype == dtype, (k, ref_sd[k].shape, shape)
    # deterministic buffers must equal what the reference constructor computes
    for k in ref_sd:
        if k.endswith(".grid") or k.endswith(".pos_offset"):
            assert torch.equal(ref_sd[k], sd[k]), k
    model.load_state_dict(sd, strict=True)
    return tc, mdl, spec, sd, model


def f32(t):
    return t.detach().to(torch.float32).cpu().numpy().copy()


def case_tiny(ref, out):
    small_vit_patch(ref, layers=2, image=32)
    over = dict(vit_layers=2, vit_image=32)
    tc, mdl, spec, sd, model = build(ref, "tiny", over)
    model.eval()
    B, S = 3, 20
    images = synth_images(B, 32, seed=11)
    labels = synth_labels(B, S, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=spec["vocab_size"] - 1)
    ids = torch.where(labels != -100, labels, torch.full_like(labels, spec["vocab_size"] - 1))
    msk = labels != -100
    with torch.no_grad():
        o1 = model(images=images, ids=ids, attn_msk=msk)
        o2 = model(images=images, ids=ids, attn_msk=None)
        m2d = torch.rand(S, S, generator=torch.Generator().manual_seed(5)) > 0.3
        o3 = model(images=images, ids=ids, attn_msk=m2d)
    out["tiny_fwd"] = dict(labels=labels.numpy(), enc=f32(o1.encoder_output), logits_rowmask=f32(o1.logits),
                           hidden_rowmask=f32(o1.hidden_state), logits_nomask=f32(o2.logits),
                           logits_mask2d=f32(o3.logits), mask2d=m2d.numpy())
    # sampling
    prompt1 = torch.full((B, 1), spec["vocab_size"] - 1, dtype=torch.long)
    prompt4 = torch.randint(0, spec["vocab_size"], (B, 4), generator=torch.Generator().manual_seed(6))
    g = {}
    g["greedy_p1"] = model.generate(images, prompt1, max_new_tokens=24, top_k=1).numpy()
    g["greedy_p4"] = model.generate(images, prompt4, max_new_tokens=16, top_k=1).numpy()
    g["prompt4"] = prompt4.numpy()
    torch.manual_seed(1234)
    g["topk5_seed1234"] = model.generate(images, prompt1, max_new_tokens=16, temperature=0.8, top_k=5).numpy()
    torch.manual_seed(4321)
    g["nucleus_seed4321"] = model.generate(images, prompt1, max_new_tokens=16, temperature=0.7, nucleus_p=0.6).numpy()
    torch.manual_seed(99)
    g["plain_seed99"] = model.generate(images, prompt1, max_new_tokens=8).numpy()
    out["tiny_generate"] = g
    # training step: plain and (moco + inverse_sqrt + eos weight)
    tok = ref_harness.fake_tokenizer(vocab_size=spec["vocab_size"], eos=spec["vocab_size"] - 1,
                                     bos=spec["vocab_size"] - 1)
    tr = {}
    for name, tcfg in (("plain", {}), ("moco", dict(moco_momentum=0.9, moco_alpha=0.4, weight_fn="inverse_sqrt_position",
                                                    eos_token_weight=2.0, training_temperature=1.3))):
        cfg = ref.CM.VisionEncoderDecoderConfig.model_validate(mdl)
        wrapper = ref.TW.ModelTrainerWrapper(cfg, tok, ref.CT.TrainerWrapperConfig(**tcfg), -100)
        wrapper.model.load_state_dict(sd, strict=True)
        if wrapper.model_m is not None:
            sd_m = synth_state_dict(spec, seed=7)
            wrapper.model_m.load_state_dict(sd_m, strict=True)
        wrapper.train()
        loss, metrics = wrapper.train_step(images, labels)
        loss.backward()
        tr[f"{name}_loss"] = f32(loss)
        named = dict(wrapper.model.named_parameters())
        for k, p in named.items():
            if p.grad is not None:
                tr[f"{name}_gnorm::{k}"] = f32(p.grad.norm())
        for k in ("decoder.transformer.h.0.cross_attn.in_proj_weight", "decoder.transformer.h.3.mlp.c_fc.bias",
                  "decoder.transformer.wpe.weight", "decoder.transformer.h.2.ln_3.weight",
                  "decoder.transformer.h.1.attn.c_attn.weight", "encoder.lsh_emb.1.emb.0.emb.weight"):
            tr[f"{name}_grad::{k}"] = f32(named[k].grad)
        if wrapper.model_m is not None:
            pm = dict(wrapper.model_m.named_parameters())
            for k in ("decoder.transformer.h.0.attn.c_attn.weight", "encoder.model.encoder.ln.bias"):
                tr[f"{name}_ema::{k}"] = f32(pm[k])
        with torch.no_grad():
            wrapper.eval()
            vloss, _ = wrapper.val_step(images, labels)
            tr[f"{name}_val_loss"] = f32(vloss)
    out["tiny_train"] = tr
    restore_vit_patch(ref)


def case_ngram(ref, out):
    from transformers import LogitsProcessorList, NoRepeatNGramLogitsProcessor
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 6, (5, 37), generator=g)          # tiny alphabet -> many repeated n-grams
    scores = torch.randn(5, 16, generator=g)
    proc = LogitsProcessorList([NoRepeatNGramLogitsProcessor(n) for n in (2, 3, 4, 5)])
    res = {"ids": ids.numpy(), "scores": scores.numpy(), "banned_full": np.isinf(proc(ids, scores).numpy())}
    for L in (1, 2, 3, 4, 5, 9):
        res[f"banned_len{L}"] = np.isinf(proc(ids[:, :L], scores).numpy())
    out["ngram"] = res


def case_optim(ref, out):
    g = torch.Generator().manual_seed(8)
    res = {}
    for name, mk in (("adamw", lambda ps: torch.optim.AdamW([dict(params=ps, lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1)])),
                     ("adamw_nowd", lambda ps: torch.optim.AdamW([dict(params=ps, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0)])),
                     ("snradam", lambda ps: ref.OPT.SNRAdam([dict(params=ps, lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1)]))):
        p = torch.nn.Parameter(torch.randn(61, 33, generator=g))
        res[f"{name}_p0"] = f32(p)
        opt = mk([p])
        for step in range(4):
            grad = torch.randn(61, 33, generator=g) * (0.5 + step)
            res[f"{name}_g{step}"] = f32(grad)
            p.grad = grad.clone()
            opt.step()
            res[f"{name}_p{step + 1}"] = f32(p)
    out["optim"] = res


def case_nano(ref, out):
    tc, mdl, spec, sd, model = build(ref, "nano", {})
    model.eval()
    B, S = 2, 24
    images = synth_images(B, 224, seed=21)
    labels = synth_labels(B, S, spec["vocab_size"], seed=22, min_len=6, max_len=20)
    ids = torch.where(labels != -100, labels, torch.full_like(labels, 50256))
    msk = labels != -100
    t0 = time.time()
    with torch.no_grad():
        o = model(images=images, ids=ids, attn_msk=msk)
    print("nano fwd", time.time() - t0)
    lg = o.logits
    out["nano_fwd"] = dict(labels=labels.numpy(), enc=f32(o.encoder_output), logits_head=f32(lg[..., :256]),
                           logits_tail=f32(lg[..., -64:]), logits_lse=f32(torch.logsumexp(lg, -1)),
                           logits_argmax=lg.argmax(-1).numpy(), hidden=f32(o.hidden_state),
                           logits_absmax=f32(lg.abs().amax()))
    # the bench workload: 8 captions x 64 new tokens, greedy (top_k = 1), prompt = [[50256]]
    N = 8
    gi = synth_images(N, 224, seed=1234)
    prompt = torch.full((N, 1), 50256, dtype=torch.long)
    t0 = time.time()
    gen = model.generate(gi, prompt, max_new_tokens=64, temperature=1.0, top_k=1)
    dt = time.time() - t0
    print("nano generate 8x64", dt, "s ->", N * 64 / dt, "tok/s")
    out["nano_generate"] = dict(greedy=gen.numpy(), seconds=np.float64(dt), threads=np.int64(torch.get_num_threads()))
    # training loss on the nano config (dropout 0), weights as above
    tok = ref_harness.fake_tokenizer()
    cfg = ref.CM.VisionEncoderDecoderConfig.model_validate(mdl)
    wrapper = ref.TW.ModelTrainerWrapper(cfg, tok, ref.CT.TrainerWrapperConfig(), -100)
    wrapper.model.load_state_dict(sd, strict=True)
    wrapper.train()
    lab = synth_labels(2, 256, spec["vocab_size"], seed=23)
    loss, _ = wrapper.train_step(images, lab)
    loss.backward()
    tr = {"labels": lab.numpy(), "loss": f32(loss)}
    named = dict(wrapper.model.named_parameters())
    for k in ("decoder.transformer.wpe.weight", "decoder.transformer.h.0.ln_3.weight",
              "decoder.transformer.h.10.cross_attn.out_proj.bias", "decoder.transformer.h.4.cross_attn.in_proj_bias"):
        tr[f"grad::{k}"] = f32(named[k].grad)
    for k, p in named.items():
        if p.grad is not None and (".h.0." in k or ".h.11." in k or "lsh_emb.0." in k):
            tr[f"gnorm::{k}"] = f32(p.grad.norm())
    out["nano_train"] = tr


def case_gpt2(ref, out):
    tc, mdl, spec, sd, model = build(ref, "gpt2", {})
    model.eval()
    B, S = 2, 16
    images = synth_images(B, 224, seed=31)
    labels = synth_labels(B, S, 50257, seed=32, min_len=5, max_len=12)
    ids = torch.where(labels != -100, labels, torch.full_like(labels, 50256))
    with torch.no_grad():
        o = model(images=images, ids=ids, attn_msk=labels != -100)
    lg = o.logits
    res = dict(labels=labels.numpy(), enc=f32(o.encoder_output), logits_head=f32(lg[..., :256]),
               logits_lse=f32(torch.logsumexp(lg, -1)), logits_argmax=lg.argmax(-1).numpy(), hidden=f32(o.hidden_state))
    prompt = torch.full((B, 1), 50256, dtype=torch.long)
    res["greedy"] = model.generate(images, prompt, max_new_tokens=12, top_k=1).numpy()
    out["gpt2_fwd"] = res


def case_beam(ref, out):
    """reference models/generation_utils.BeamSearchTokenGenerator on the tiny model, deterministic settings
    (temperature 0 -> arg-top expansions, consolidation_temperature 0 -> top beams)."""
    small_vit_patch(ref, layers=2, image=32)
    over = dict(vit_layers=2, vit_image=32)
    tc, mdl, spec, sd, model = build(ref, "tiny", over)
    model.eval()
    from models.generation_utils import BeamSearchTokenGenerator
    images = synth_images(2, 32, seed=11)
    eos = spec["vocab_size"] - 1
    prompt = torch.full((2, 1), eos, dtype=torch.long)
    g = {}
    for name, kw in (("plain", dict(beam_width=3, temperature=0.0, top_k=None, max_new_tokens=10, beam_expansion_factor=4,
                                    eos_token_id=611, consolidation_temperature=0.0, length_boost=1.0)),
                     ("topk_eos", dict(beam_width=4, temperature=0.0, top_k=12, max_new_tokens=12, beam_expansion_factor=3,
                                       eos_token_id=7, consolidation_temperature=0.0, length_boost=1.3))):
        gen = BeamSearchTokenGenerator(model, **kw)
        with torch.no_grad():
            ids, scores = gen(images, prompt)
        g[name + "_ids"] = ids.numpy()
        g[name + "_scores"] = f32(scores)
    out["tiny_beam"] = g
    restore_vit_patch(ref)


def case_contrastive(ref, out):
    """training/wrapper.py:98-118,206-209: the contrastive auxiliary loss (add_contrastive_loss), tiny model."""
    small_vit_patch(ref, layers=2, image=32)
    over = dict(vit_layers=2, vit_image=32)
    tc, mdl, spec, sd, model = build(ref, "tiny", over)
    eos = spec["vocab_size"] - 1
    images = synth_images(3, 32, seed=11)
    labels = synth_labels(3, 20, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=eos)
    tok = ref_harness.fake_tokenizer(vocab_size=spec["vocab_size"], eos=eos, bos=eos)
    cfg = ref.CM.VisionEncoderDecoderConfig.model_validate(mdl)
    tcfg = dict(add_contrastive_loss=True, training_contrastive_temperature=0.7, weight_fn="inverse_sqrt_position", eos_token_weight=2.0)
    wrapper = ref.TW.ModelTrainerWrapper(cfg, tok, ref.CT.TrainerWrapperConfig(**tcfg), -100)
    wrapper.model.load_state_dict(sd, strict=True)
    wrapper.train()
    loss, metrics = wrapper.train_step(images, labels)
    loss.backward()
    res = dict(labels=labels.numpy(), loss=f32(loss), loss_lm=f32(metrics["train_loss_lm"]),
               loss_contrastive=f32(metrics["train_loss_contrastive"]))
    named = dict(wrapper.model.named_parameters())
    for k, p in named.items():
        if p.grad is not None:
            res[f"gnorm::{k}"] = f32(p.grad.norm())
    for k in ("decoder.transformer.wte.weight", "decoder.transformer.h.3.mlp.c_proj.weight", "decoder.transformer.ln_f.weight"):
        res[f"grad::{k}"] = f32(named[k].grad)
    out["tiny_contrastive"] = res
    restore_vit_patch(ref)


def case_peer(ref, out):
    """SURVEY.md 8f-2 (the gpu/nano.yaml variant): PretrainedViT + PEER tail + bridging Linear + cross-attention-only decoder,
    small dims (configs/tiny_peer.yaml): forward, greedy ids and one training step (loss + gradients of the PEER parameters)."""
    small_vit_patch(ref, layers=2, image=32)
    over = dict(vit_layers=2, vit_image=32)
    tc, mdl, spec, sd, model = build(ref, "tiny_peer", over)
    model.eval()
    B, S = 3, 20
    images = synth_images(B, 32, seed=11)
    labels = synth_labels(B, S, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=spec["vocab_size"] - 1)
    ids = torch.where(labels != -100, labels, torch.full_like(labels, spec["vocab_size"] - 1))
    with torch.no_grad():
        o = model(images=images, ids=ids, attn_msk=None)
    res = dict(labels=labels.numpy(), enc=f32(o.encoder_output), logits=f32(o.logits), hidden=f32(o.hidden_state))
    prompt = torch.full((B, 1), spec["vocab_size"] - 1, dtype=torch.long)
    res["greedy"] = model.generate(images, prompt, max_new_tokens=16, top_k=1).numpy()
    tok = ref_harness.fake_tokenizer(vocab_size=spec["vocab_size"], eos=spec["vocab_size"] - 1, bos=spec["vocab_size"] - 1)
    cfg = ref.CM.VisionEncoderDecoderConfig.model_validate(mdl)
    wrapper = ref.TW.ModelTrainerWrapper(cfg, tok, ref.CT.TrainerWrapperConfig(), -100)
    wrapper.model.load_state_dict(sd, strict=True)
    wrapper.train()
    loss, _ = wrapper.train_step(images, labels)
    loss.backward()
    res["train_loss"] = f32(loss)
    named = dict(wrapper.model.named_parameters())
    for k, p in named.items():
        if p.grad is not None and ("peer" in k or k == "encoder.1.weight"):
            res[f"gnorm::{k}"] = f32(p.grad.norm())
    for k in ("encoder.0.peer.query_left.linear.weight", "encoder.0.peer.query_linear.weight", "encoder.0.peer.residual.weight",
              "encoder.0.peer.emb_out.weight", "encoder.0.peer.emb_in.weight", "encoder.1.weight"):
        res[f"grad::{k}"] = f32(named[k].grad)
    res["grad_head::encoder.0.peer.key_linear.weight"] = f32(named["encoder.0.peer.key_linear.weight"].grad[:64])
    res["grad_head::encoder.0.peer_proj_wt"] = f32(named["encoder.0.peer_proj_wt"].grad[:32, :32])
    out["tiny_peer"] = res
    restore_vit_patch(ref)


def main():
    torch.manual_seed(0)
    ref = ref_harness.load_reference()
    which = sys.argv[1:] or ["tiny", "ngram", "optim", "nano", "gpt2"]
    out = {}
    for name in which:
        t0 = time.time()
        globals()["case_" + name](ref, out)
        print(f"case {name}: {time.time() - t0:.1f}s")
    for name, d in out.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
