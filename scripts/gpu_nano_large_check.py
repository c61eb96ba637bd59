"""gpu/nano.yaml at its REAL dimensions (PEER 65 536 experts, 1600 -> 1280 bridge, 36 x 1280 x 20 decoder; 1.5 G parameters):
construct, forward against the CPU oracle (fp32, 1e-4), greedy generate against the oracle's cache-less loop (bit-exact),
one SNRAdam training step, bf16 generate + tok/s.  python scripts/gpu_nano_large_check.py"""
import fnmatch
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.config_schema import TrainerWrapperConfig  # noqa: E402
from image2text_b200.optimizer import SNRAdam  # noqa: E402
from image2text_b200.synthetic import synth_images, synth_labels  # noqa: E402
from image2text_b200.wrapper import ModelTrainerWrapper  # noqa: E402
from oracle import i2t_oracle as O  # noqa: E402

tc = load_training_config(os.path.join(ROOT, "configs", "gpu_nano.yaml"))
t0 = time.time()
m = VisionEncoderDecoder(tc.model, device="cuda", seed=0, spec_overrides=dict(dropout=0.0, attn_dropout=0.0))
m.eval()
spec = m.spec
print("built %.1f M parameters in %.0f s" % (sum(p.numel() for p in m.parameters()) / 1e6, time.time() - t0), flush=True)
sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
images = synth_images(2, 224, seed=5)
ids = torch.randint(0, 50256, (2, 12), generator=torch.Generator().manual_seed(1))
with torch.no_grad():
    out = m(images=images.cuda(), ids=ids.cuda())
    enc, logits, hidden = O.ved_forward(sd, spec, images, ids, normalize_grads=False)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
print("fp32 forward vs oracle: enc %.2e hidden %.2e logits %.2e" % (rel(out.encoder_output.cpu(), enc), rel(out.hidden_state.cpu(), hidden),
                                                                   rel(out.logits.cpu(), logits)), flush=True)
assert rel(out.encoder_output.cpu(), enc) < 1e-4 and rel(out.logits.cpu(), logits) < 1e-4
prompt = torch.full((2, 1), 50256, dtype=torch.long)
got = m.generate(images.cuda(), prompt.cuda(), max_new_tokens=8, top_k=1).cpu()
want = O.generate(sd, spec, images, prompt, 8, top_k=1)
print("fp32 greedy ids equal the oracle's:", bool(torch.equal(got, want)), got[0].tolist(), flush=True)
assert torch.equal(got, want)
eng = next(iter(m._decode_engines.values()))
print("fp32 decode mode:", eng.mode)
del m
torch.cuda.empty_cache()
# one SNRAdam step on the YAML's groups (batch 4 here)
tok = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, mask_token_id=None, vocab_size=50257)
w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(), -100, device="cuda", compute_dtype=torch.bfloat16, seed=0)
w.train()
groups = []
for oc in tc.optimizers:
    ps = [p for n, p in w.named_parameters() if any(fnmatch.fnmatch(n.split(".", 1)[-1], pat) for pat in oc.target_modules)]
    groups.append(dict(params=ps, lr=oc.lr, weight_decay=oc.weight_decay, betas=oc.betas))
chosen = {id(p) for g in groups for p in g["params"]}
for p in w.model.parameters():
    if id(p) not in chosen:
        p.requires_grad_(False)
opt = SNRAdam(groups)
timg, tlab = synth_images(4, 224, seed=7).cuda(), synth_labels(4, 256, seed=7).cuda()
losses = []
for _ in range(4):
    loss, _ = w.train_step(timg, tlab)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    losses.append(float(loss.detach()))
print("bf16 SNRAdam steps on %d tensors / %.1f M parameters, loss %s" % (len(chosen), sum(p.numel() for g in groups for p in g["params"]) / 1e6,
                                                                        " ".join("%.3f" % l for l in losses)), flush=True)
assert losses[-1] < losses[0]
w.eval()
mb = w.model
img8 = synth_images(8, 224, seed=9).cuda()
p8 = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(2):
    ids8 = mb.generate(img8, p8, max_new_tokens=32, top_k=1)
torch.cuda.synchronize()
t0 = time.time()
ids8 = mb.generate(img8, p8, max_new_tokens=32, top_k=1)
torch.cuda.synchronize()
dt = time.time() - t0
eng = next(iter(mb._decode_engines.values()))
print("bf16 generate 8 x 32: %.1f ms -> %.0f tok/s (decode mode %s)" % (dt * 1e3, 8 * 32 / dt, eng.mode))
print("ok")
