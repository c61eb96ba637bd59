"""Pin the CPU oracle (oracle/i2t_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from image2text_b200.model_spec import synth_state_dict
from image2text_b200.synthetic import synth_images, synth_labels
from oracle import i2t_oracle as O
from tests.helpers import rel_err, spec_and_weights

TOL = 2e-5  # CPU fp32 vs CPU fp32, different op order only


def T(x):
    return torch.from_numpy(np.asarray(x))


def test_tiny_forward_all_mask_kinds(golden):
    g = golden("tiny_fwd")
    _, spec, sd = spec_and_weights("tiny")
    images = synth_images(3, 32, seed=11)
    labels = T(g["labels"])
    eos = spec["vocab_size"] - 1
    ids = torch.where(labels != -100, labels, torch.full_like(labels, eos))
    with torch.no_grad():
        enc, logits, hidden = O.ved_forward(sd, spec, images, ids, attn_msk=labels != -100)
        assert rel_err(enc, T(g["enc"])) < TOL
        assert rel_err(logits, T(g["logits_rowmask"])) < TOL
        assert rel_err(hidden, T(g["hidden_rowmask"])) < TOL
        _, logits, _ = O.ved_forward(sd, spec, images, ids, attn_msk=None)
        assert rel_err(logits, T(g["logits_nomask"])) < TOL
        _, logits, _ = O.ved_forward(sd, spec, images, ids, attn_msk=T(g["mask2d"]))
        assert rel_err(logits, T(g["logits_mask2d"])) < TOL


def test_tiny_generate_matches_reference_ids(golden):
    g = golden("tiny_generate")
    _, spec, sd = spec_and_weights("tiny")
    images = synth_images(3, 32, seed=11)
    eos = spec["vocab_size"] - 1
    p1 = torch.full((3, 1), eos, dtype=torch.long)
    assert np.array_equal(O.generate(sd, spec, images, p1, 24, top_k=1).numpy(), g["greedy_p1"])
    assert np.array_equal(O.generate(sd, spec, images, T(g["prompt4"]), 16, top_k=1).numpy(), g["greedy_p4"])
    # same torch RNG stream, same number of multinomial draws -> identical samples
    torch.manual_seed(1234)
    assert np.array_equal(O.generate(sd, spec, images, p1, 16, temperature=0.8, top_k=5).numpy(), g["topk5_seed1234"])
    torch.manual_seed(4321)
    assert np.array_equal(O.generate(sd, spec, images, p1, 16, temperature=0.7, nucleus_p=0.6).numpy(),
                          g["nucleus_seed4321"])
    torch.manual_seed(99)
    assert np.array_equal(O.generate(sd, spec, images, p1, 8).numpy(), g["plain_seed99"])


def test_ngram_ban_known_answers(golden):
    g = golden("ngram")
    ids, scores = T(g["ids"]), T(g["scores"])
    got = torch.isinf(O.apply_ngram_ban(ids, scores, (2, 3, 4, 5))).numpy()
    assert np.array_equal(got, g["banned_full"])
    assert g["banned_full"].any()
    for L in (1, 2, 3, 4, 5, 9):
        got = torch.isinf(O.apply_ngram_ban(ids[:, :L], scores, (2, 3, 4, 5))).numpy()
        assert np.array_equal(got, g[f"banned_len{L}"])


@pytest.mark.parametrize("name", ["plain", "moco"])
def test_tiny_train_step_loss_and_grads(golden, name):
    g = golden("tiny_train")
    _, spec, sd0 = spec_and_weights("tiny")
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd0.items()}
    sd["decoder.lm_head.weight"] = sd["decoder.transformer.wte.weight"]
    images = synth_images(3, 32, seed=11)
    eos = spec["vocab_size"] - 1
    labels = synth_labels(3, 20, spec["vocab_size"], seed=12, min_len=3, max_len=14, eos=eos)
    kw = {}
    sd_m = None
    if name == "moco":
        sd_m = synth_state_dict(spec, seed=7)
        kw = dict(alpha=0.4, temperature=1.3, weight_fn="inverse_sqrt_position", eos_token_weight=2.0, eos_token_id=eos)
    ids, msk = O.wrapper_inputs(labels, eos_token_id=eos, bos_token_id=eos)
    _, logits, _ = O.ved_forward(sd, spec, images, ids, attn_msk=msk)
    logits_m = None
    if sd_m is not None:
        with torch.no_grad():
            _, logits_m, _ = O.ved_forward(sd_m, spec, images, ids, attn_msk=msk)
    loss = O.lm_loss(logits, labels, logits_m, **kw)
    assert abs(float(loss.detach()) - float(g[f"{name}_loss"])) < 1e-5 * abs(float(g[f"{name}_loss"]))
    # val_step never uses the teacher (training/wrapper.py:200-203); dropout is 0 so plain train == val
    kw_val = {k: v for k, v in kw.items() if k != "alpha"}
    val = O.lm_loss(logits.detach(), labels, None, **kw_val)
    assert abs(float(val) - float(g[f"{name}_val_loss"])) < 1e-5 * abs(float(val))
    loss.backward()
    checked = 0
    for key, val in g.items():
        if key.startswith(f"{name}_gnorm::"):
            k = key.split("::")[1]
            if k == "decoder.lm_head.weight":
                k = "decoder.transformer.wte.weight"
            gr = sd[k].grad
            gn = 0.0 if gr is None else float(gr.norm())
            assert abs(gn - float(val)) <= 1e-4 * max(float(val), 1e-8), (k, gn, float(val))
            checked += 1
        if key.startswith(f"{name}_grad::"):
            k = key.split("::")[1]
            assert rel_err(sd[k].grad, T(val)) < 1e-4, k
    assert checked > 50
    if name == "moco":
        pm = [sd_m["decoder.transformer.h.0.attn.c_attn.weight"].clone(), sd_m["encoder.model.encoder.ln.bias"].clone()]
        O.ema_update(pm, [sd0["decoder.transformer.h.0.attn.c_attn.weight"], sd0["encoder.model.encoder.ln.bias"]], 0.9)
        assert rel_err(pm[0], T(g["moco_ema::decoder.transformer.h.0.attn.c_attn.weight"])) < 1e-6
        assert rel_err(pm[1], T(g["moco_ema::encoder.model.encoder.ln.bias"])) < 1e-6


@pytest.mark.parametrize("name,fn,kw", [
    ("adamw", "adamw_step", dict(lr=3e-3, beta1=0.9, beta2=0.95, weight_decay=0.1)),
    ("adamw_nowd", "adamw_step", dict(lr=1e-3, beta1=0.9, beta2=0.999, weight_decay=0.0)),
    ("snradam", "snradam_step", dict(lr=3e-3, beta1=0.9, beta2=0.95, weight_decay=0.1)),
])
def test_optimizer_steps(golden, name, fn, kw):
    g = golden("optim")
    p = T(g[f"{name}_p0"]).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(4):
        getattr(O, fn)(p, T(g[f"{name}_g{step}"]), m, v, step + 1, **kw)
        assert rel_err(p, T(g[f"{name}_p{step + 1}"])) < 2e-6, step


def test_nano_forward(golden):
    g = golden("nano_fwd")
    _, spec, sd = spec_and_weights("nano")
    images = synth_images(2, 224, seed=21)
    labels = T(g["labels"])
    ids = torch.where(labels != -100, labels, torch.full_like(labels, 50256))
    with torch.no_grad():
        enc, logits, hidden = O.ved_forward(sd, spec, images, ids, attn_msk=labels != -100)
    assert rel_err(enc, T(g["enc"])) < TOL
    assert rel_err(hidden, T(g["hidden"])) < 5e-5
    scale = float(g["logits_absmax"])
    assert float((logits[..., :256] - T(g["logits_head"])).abs().max()) < 5e-5 * scale
    assert float((logits[..., -64:] - T(g["logits_tail"])).abs().max()) < 5e-5 * scale
    assert float((torch.logsumexp(logits, -1) - T(g["logits_lse"])).abs().max()) < 1e-4
    assert np.array_equal(logits.argmax(-1).numpy(), g["logits_argmax"])


def test_nano_greedy_bench_workload(golden):
    """8 captions x 64 new tokens, top_k=1: the oracle reproduces the reference's ids bit-exactly.
    (16 tokens here to keep the CPU suite short; the GPU test checks all 64.)"""
    g = golden("nano_generate")
    _, spec, sd = spec_and_weights("nano")
    images = synth_images(8, 224, seed=1234)
    prompt = torch.full((8, 1), 50256, dtype=torch.long)
    ids = O.generate(sd, spec, images, prompt, 16, top_k=1)
    assert np.array_equal(ids.numpy(), g["greedy"][:, :17])


def test_nano_train_loss_and_grads(golden):
    g = golden("nano_train")
    _, spec, sd0 = spec_and_weights("nano")
    want = ["decoder.transformer.wpe.weight", "decoder.transformer.h.0.ln_3.weight",
            "decoder.transformer.h.10.cross_attn.out_proj.bias", "decoder.transformer.h.4.cross_attn.in_proj_bias"]
    sd = dict(sd0)
    for k in want:
        sd[k] = sd0[k].clone().requires_grad_(True)
    images = synth_images(2, 224, seed=21)
    labels = T(g["labels"])
    loss = O.train_step_loss(sd, spec, images, labels)
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    loss.backward()
    for k in want:
        assert rel_err(sd[k].grad, T(g[f"grad::{k}"])) < 2e-4, k


def test_gpt2_hf_layout_forward_and_greedy(golden):
    g = golden("gpt2_fwd")
    _, spec, sd = spec_and_weights("gpt2")
    images = synth_images(2, 224, seed=31)
    labels = T(g["labels"])
    ids = torch.where(labels != -100, labels, torch.full_like(labels, 50256))
    with torch.no_grad():
        enc, logits, hidden = O.ved_forward(sd, spec, images, ids, attn_msk=labels != -100)
    assert rel_err(enc, T(g["enc"])) < TOL
    assert rel_err(hidden, T(g["hidden"])) < 5e-5
    assert rel_err(logits[..., :256], T(g["logits_head"])) < 5e-5
    assert float((torch.logsumexp(logits, -1) - T(g["logits_lse"])).abs().max()) < 1e-4
    prompt = torch.full((2, 1), 50256, dtype=torch.long)
    got = O.generate(sd, spec, images, prompt, 12, top_k=1)
    assert np.array_equal(got.numpy(), g["greedy"])


def test_peer_tail_variant_forward_greedy_and_grads(golden):
    """SURVEY.md 8f-2 (gpu/nano.yaml variant): PEER tail + bridging Linear + cross-attention-only decoder (configs/tiny_peer.yaml)
    against the reference-made fixture: encoder output, logits, greedy ids, train-step loss and the PEER gradients."""
    g = golden("tiny_peer")
    _, spec, sd0 = spec_and_weights("tiny_peer")
    assert spec["tail"] == "peer" and not spec["use_soft_prompting"]
    images = synth_images(3, 32, seed=11)
    labels = T(g["labels"])
    eos = spec["vocab_size"] - 1
    ids = torch.where(labels != -100, labels, torch.full_like(labels, eos))
    with torch.no_grad():
        enc, logits, hidden = O.ved_forward(sd0, spec, images, ids)
    assert rel_err(enc, T(g["enc"])) < TOL
    assert rel_err(logits, T(g["logits"])) < TOL
    assert rel_err(hidden, T(g["hidden"])) < TOL
    prompt = torch.full((3, 1), eos, dtype=torch.long)
    assert np.array_equal(O.generate(sd0, spec, images, prompt, 16, top_k=1).numpy(), g["greedy"])
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd0.items()}
    sd["decoder.lm_head.weight"] = sd["decoder.transformer.wte.weight"]
    tids, msk = O.wrapper_inputs(labels, eos_token_id=eos, bos_token_id=eos)
    _, lg, _ = O.ved_forward(sd, spec, images, tids, attn_msk=msk)
    loss = O.lm_loss(lg, labels, None)
    assert abs(float(loss.detach()) - float(g["train_loss"])) < 1e-5 * abs(float(g["train_loss"]))
    loss.backward()
    n = 0
    for key, val in g.items():
        if key.startswith("gnorm::"):
            k = key.split("::")[1]
            assert abs(float(sd[k].grad.norm()) - float(val)) <= 1e-4 * max(float(val), 1e-8), k
            n += 1
        elif key.startswith("grad::"):
            assert rel_err(sd[key.split("::")[1]].grad, T(val)) < 1e-4, key
        elif key.startswith("grad_head::"):
            k = key.split("::")[1]
            gr = sd[k].grad
            want = T(val)
            got = gr[:want.shape[0]] if want.dim() == 2 and gr.dim() == 2 else gr[:32, :32]
            assert rel_err(got, want) < 1e-4, key
    assert n >= 9


def test_contrastive_auxiliary_loss(golden):
    """training/wrapper.py:98-118,206-209 (add_contrastive_loss) against the reference-made fixture: both losses and gradients."""
    g = golden("tiny_contrastive")
    _, spec, sd0 = spec_and_weights("tiny")
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd0.items()}
    sd["decoder.lm_head.weight"] = sd["decoder.transformer.wte.weight"]
    images = synth_images(3, 32, seed=11)
    eos = spec["vocab_size"] - 1
    labels = T(g["labels"])
    ids, msk = O.wrapper_inputs(labels, eos_token_id=eos, bos_token_id=eos)
    _, logits, hidden = O.ved_forward(sd, spec, images, ids, attn_msk=msk)
    wkw = dict(weight_fn="inverse_sqrt_position", eos_token_weight=2.0, eos_token_id=eos)
    lm = O.lm_loss(logits, labels, None, **wkw)
    con = O.contrastive_loss(sd, hidden, labels, temperature=0.7, **wkw)
    assert abs(float(lm.detach()) - float(g["loss_lm"])) < 1e-5 * abs(float(g["loss_lm"]))
    assert abs(float(con.detach()) - float(g["loss_contrastive"])) < 1e-5 * abs(float(g["loss_contrastive"]))
    (lm + con).backward()
    n = 0
    for key, val in g.items():
        if key.startswith("gnorm::"):
            k = key.split("::")[1]
            if k == "decoder.lm_head.weight":
                k = "decoder.transformer.wte.weight"
            assert abs(float(sd[k].grad.norm()) - float(val)) <= 1e-4 * max(float(val), 1e-8), k
            n += 1
        elif key.startswith("grad::"):
            assert rel_err(sd[key.split("::")[1]].grad, T(val)) < 1e-4, key
    assert n > 50
