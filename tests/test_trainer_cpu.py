"""Host logic of the accelerate-free launcher (image2text_b200/trainer.py) against what the reference's trainer does
(trainer.py:145-172 parameter groups, training/utils.py:104-123 partial checkpoints, models/utils.py:31-36 resume); CPU
only -- no kernel runs."""
import fnmatch
import os
import types

import torch

from image2text_b200 import load_training_config
from image2text_b200.config_schema import TrainerWrapperConfig
from image2text_b200.trainer import PatternMatcher, build_param_groups, checkpoint_state, save_checkpoint, synthetic_batches
from image2text_b200.wrapper import ModelTrainerWrapper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wrapper(name, seed=3, **trainer_kw):
    tc = load_training_config(os.path.join(ROOT, "configs", name + ".yaml"))
    tok = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, mask_token_id=None, vocab_size=50257)
    over = dict(vit_layers=1, vit_image=32) if name == "tiny" else {}
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(**trainer_kw), -100, device="cpu", spec_overrides=over, seed=seed)
    return tc, w


def test_pattern_matcher_is_fnmatch_any():
    m = PatternMatcher(["decoder*.transformer.h.*.cross_attn.*", "encoder*.lsh_emb.*"])
    assert m.match("decoder.transformer.h.10.cross_attn.in_proj_weight")
    assert m.match("encoder.lsh_emb.3.emb.1.emb.weight")
    assert not m.match("decoder.transformer.h.1.attn.c_attn.weight") and not m.match("model.decoder.transformer.wpe.weight")


def test_nano_param_groups_and_partial_checkpoint(tmp_path):
    """SURVEY Q5 (probed on the reference): nano.yaml optimises 61 tensors / 21.3 M of 230.7 M parameters, and its
    checkpoint holds exactly those keys."""
    tc, w = _wrapper("nano", moco_momentum=0.995, moco_alpha=0.4)
    groups, matchers = build_param_groups(w, tc)
    assert [len(g["params"]) for g in groups] == [25, 36]
    assert round(sum(p.numel() for g in groups for p in g["params"]) / 1e6, 1) == 21.3
    assert round(sum(p.numel() for p in w.model.parameters()) / 1e6, 1) == 230.7
    assert groups[0]["lr"] == 1e-3 and groups[1]["lr"] == 6e-4 and groups[0]["betas"] == (0.9, 0.95)
    teacher = {id(p) for p in w.model_m.parameters()}
    assert not any(id(p) in teacher for g in groups for p in g["params"])          # model_m is never optimised
    # names: the reference strips the leading "model." before matching
    pats = [p for oc in tc.optimizers for p in oc.target_modules]
    want = [n for n, _ in w.model.named_parameters() if any(fnmatch.fnmatch(n, p) for p in pats)]
    state = checkpoint_state(w.model, matchers)
    assert list(state) == want and len(want) == 61
    path = str(tmp_path / "nano_partial.pt")
    save_checkpoint(w.model, path, matchers)
    on_disk = torch.load(path)
    assert list(on_disk) == want
    # resume into a differently initialised model: matched tensors come from the file, everything else is untouched
    _, w2 = _wrapper("nano", seed=4)
    before = {k: v.clone() for k, v in w2.model.state_dict().items()}
    for k in want[:3]:
        assert not torch.equal(before[k], on_disk[k])
    w2.model.load_partial_checkpoint(path)
    after = w2.model.state_dict()
    for k, v in after.items():
        assert torch.equal(v, on_disk[k] if k in on_disk else before[k]), k


def test_single_group_without_globs_saves_everything(tmp_path):
    tc, w = _wrapper("tiny")
    groups, matchers = build_param_groups(w, tc)
    assert matchers == [] and len(groups) == 1
    assert len(groups[0]["params"]) == len(list(w.model.parameters()))
    assert list(checkpoint_state(w.model, matchers)) == list(w.model.state_dict())


def test_synthetic_batches_have_the_loader_shapes():
    it = synthetic_batches(4, 32, 613, 612, seed=5, width=20, pool=2)
    im, lb = next(it)
    assert im.shape == (4, 3, 32, 32) and lb.shape == (4, 20) and lb.dtype == torch.int64
    assert bool(((lb == -100) | ((lb >= 0) & (lb < 613))).all()) and bool((lb == 612).any(dim=1).all())
    im2, _ = next(it)
    im3, _ = next(it)
    assert not torch.equal(im, im2) and torch.equal(im, im3)
