"""Decode throughput of the HF GPT-2 layout decoder (configs/gpt2.yaml: ViT-B/16 trainable trunk + per-slot MLP tail + GPT-2 124M
with cross attention in every block, 16-token soft prompt): images x 64 new tokens, greedy.
    python scripts/sweep_decode_hf.py [bf16|fp32] [n_sequences,...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
points = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 64, 256]
tc = load_training_config(os.path.join(ROOT, "configs", "gpt2.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=dtype)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
for B in points:
    images = synth_images(max(1, B // 8), 224, seed=1234).cuda().repeat_interleave(8, dim=0)[:B]
    prompt = torch.full((B, 1), 50256, dtype=torch.long, device="cuda")
    for _ in range(2):
        m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=1, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=1, seed=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"decoder": "hf_gpt2", "sequences": B, "top_k": 1, "ms_per_generate": round(ms, 2),
                      "tok_per_s": round(B * 64 / (ms / 1e3), 1), "dtype": str(dtype)}), flush=True)
