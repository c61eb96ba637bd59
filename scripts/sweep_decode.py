"""Decode throughput vs. sequences per GPU (SURVEY config 2 / 5 sweep): images x 8 captions, 64 new tokens, greedy or top-k.
    python scripts/sweep_decode.py [bf16|fp32]
B <= 8 runs the one-launch megakernel; larger batches run the per-stage kernels under one CUDA graph per step."""
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
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=dtype)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
POINTS = ((1, 1), (2, 1), (4, 1), (8, 1), (8, 16), (32, 1), (64, 1), (64, 50))
if len(sys.argv) > 2:
    POINTS = tuple((int(x), 1) for x in sys.argv[2].split(","))
for n_img, top_k in POINTS:
    B = n_img * 8
    images = synth_images(n_img, 224, seed=1234).cuda().repeat_interleave(8, dim=0)
    prompt = torch.full((B, 1), 50256, dtype=torch.long, device="cuda")
    try:
        for _ in range(2):
            m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=top_k, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=top_k, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        eng = m._decode_engines[(B, dtype, False)]
        print(json.dumps({"images": n_img, "sequences": B, "top_k": top_k, "mode": eng.mode, "ms_per_generate": round(ms, 2),
                          "tok_per_s": round(B * 64 / (ms / 1e3), 1), "dtype": str(dtype)}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"images": n_img, "sequences": B, "top_k": top_k, "error": repr(e)[:300]}), flush=True)
