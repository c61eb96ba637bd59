"""Kernel-time table (torch profiler / CUPTI) of one generate() call at a given number of sequences.
    python scripts/profile_decode_batch.py <n_images> [top_k]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

n_img = int(sys.argv[1])
top_k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=torch.bfloat16)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
B = n_img * 8
images = synth_images(n_img, 224, seed=1234).cuda().repeat_interleave(8, dim=0)
prompt = torch.full((B, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(2):
    m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=top_k, seed=1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=top_k, seed=1)
    torch.cuda.synchronize()
print(f"sequences {B} top_k {top_k}")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
