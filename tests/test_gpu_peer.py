"""SURVEY.md 8f-2, the gpu/nano.yaml variant: PretrainedViT + PEER tail (reference models/layers.py:21-109, encoder.py:114-115)
+ bridging Linear (vision_encoder_decoder.py:33-37) + cross-attention-only decoder + SNRAdam, against the reference-made fixture
tests/golden/tiny_peer.npz (configs/tiny_peer.yaml: the same structure at test size) and the oracle."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from image2text_b200 import VisionEncoderDecoder  # noqa: E402
from image2text_b200.config_schema import TrainerWrapperConfig  # noqa: E402
from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.optimizer import SNRAdam  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402
from image2text_b200.wrapper import ModelTrainerWrapper  # noqa: E402
from tests.helpers import SPEC_OVERRIDES, check_picks_vs_oracle, rel_err, spec_and_weights  # noqa: E402


def T(x):
    return torch.from_numpy(np.asarray(x))


def build(dtype=torch.float32):
    tc, spec, sd = spec_and_weights("tiny_peer")
    m = VisionEncoderDecoder(tc.model, spec_overrides=SPEC_OVERRIDES["tiny_peer"], device="cuda", compute_dtype=dtype)
    m.load_state_dict(sd)
    return m.eval(), spec, sd


def test_peer_variant_forward_and_greedy_match_reference(golden):
    g = golden("tiny_peer")
    m, spec, sd = build()
    assert list(m.state_dict().keys())[0] == "encoder.0.peer_proj_wt" and "encoder.1.weight" in m.state_dict()
    images = synth_images(3, 32, seed=11).cuda()
    labels = T(g["labels"])
    eos = spec["vocab_size"] - 1
    ids = torch.where(labels != -100, labels, torch.full_like(labels, eos)).cuda()
    with torch.no_grad():
        out = m(images=images, ids=ids)
    assert rel_err(out.encoder_output.cpu(), T(g["enc"])) < 1e-4
    assert rel_err(out.hidden_state.cpu(), T(g["hidden"])) < 1e-4
    assert rel_err(out.logits.cpu(), T(g["logits"])) < 1e-4
    prompt = torch.full((3, 1), eos, dtype=torch.long, device="cuda")
    got = m.generate(images, prompt, max_new_tokens=16, top_k=1)
    assert np.array_equal(got.cpu().numpy(), g["greedy"])                      # fp32: bit-exact greedy ids (no soft prompt rows)


def test_peer_variant_train_step_matches_reference(golden):
    g = golden("tiny_peer")
    tc, spec, sd = spec_and_weights("tiny_peer")
    eos = spec["vocab_size"] - 1
    tok = types.SimpleNamespace(eos_token_id=eos, bos_token_id=eos, mask_token_id=None, vocab_size=spec["vocab_size"])
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(), -100, device="cuda", spec_overrides=SPEC_OVERRIDES["tiny_peer"])
    w.model.load_state_dict(sd)
    w.train()
    images = synth_images(3, 32, seed=11).cuda()
    labels = T(g["labels"]).cuda()
    loss, _ = w.train_step(images, labels)
    assert abs(float(loss) - float(g["train_loss"])) < 1e-4 * abs(float(g["train_loss"]))
    loss.backward()
    named = dict(w.model.named_parameters())
    n = 0
    for key, val in g.items():
        if key.startswith("gnorm::"):
            k = key.split("::")[1]
            assert abs(float(named[k].grad.norm()) - float(val)) <= 3e-4 * max(float(val), 1e-7), (k, float(named[k].grad.norm()), float(val))
            n += 1
        elif key.startswith("grad::"):
            assert rel_err(named[key.split("::")[1]].grad.cpu(), T(val)) < 3e-4, key
        elif key.startswith("grad_head::"):
            k = key.split("::")[1]
            want = T(val)
            gr = named[k].grad.cpu()
            got = gr[:want.shape[0]] if want.dim() == 2 else gr[:32, :32]
            assert rel_err(got, want) < 3e-4, key
    assert n >= 9
    # the YAML's optimiser: SNRAdam on the PEER / wpe / cross-attention groups steps without error and moves the experts
    import fnmatch
    groups = []
    for oc in tc.optimizers:
        ps = [p for nme, p in w.named_parameters() if nme.split(".", 1)[0] != "model_m" and
              any(fnmatch.fnmatch(nme.split(".", 1)[-1], pat) for pat in oc.target_modules)]
        groups.append(dict(params=ps, lr=oc.lr, weight_decay=oc.weight_decay, betas=oc.betas))
    assert any(p is named["encoder.0.peer.emb_out.weight"] for p in groups[0]["params"])
    before = named["encoder.0.peer.emb_out.weight"].detach().clone()
    SNRAdam(groups).step()
    assert not torch.equal(before, named["encoder.0.peer.emb_out.weight"].detach())


def test_peer_variant_bf16_decode_teacher_forced():
    m, spec, sd = build(torch.bfloat16)
    images = synth_images(3, 32, seed=11).cuda()
    eos = spec["vocab_size"] - 1
    prompt = torch.full((3, 1), eos, dtype=torch.long, device="cuda")
    eng = DecodeEngine(m, 3)
    assert eng.mode == "mega3" and eng.n_prompt == 0
    got = eng.generate(images, prompt, 20, 1.0, 1, seed=0)
    check_picks_vs_oracle("tiny_peer", m, images, got, 1, top_k=1)
