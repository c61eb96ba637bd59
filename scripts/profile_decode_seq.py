"""Per-launch kernel durations (CUPTI, in launch order) of one KV-cached decode step at a given number of sequences.
    python scripts/profile_decode_seq.py <n_images> [n_kernels]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

n_img = int(sys.argv[1])
n_show = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=torch.bfloat16)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
B = n_img * 8
images = synth_images(n_img, 224, seed=1234).cuda().repeat_interleave(8, dim=0)
prompt = torch.full((B, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(2):
    m.generate(images, prompt, max_new_tokens=8, temperature=1.0, top_k=1, seed=1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.generate(images, prompt, max_new_tokens=8, temperature=1.0, top_k=1, seed=1)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# the last decode step = the tail of the launch sequence; find the last dec_embed kernel
idx = max(i for i, e in enumerate(evs) if "dec_embed" in e.name)
t0 = evs[idx].time_range.start
print(f"sequences {B}: last decode step, {len(evs) - idx} launches, {evs[-1].time_range.end - t0:.1f} us")
for e in evs[idx: idx + n_show]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.2f} us  {e.name[:90]}")
print("   ...")
for e in evs[-5:]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.2f} us  {e.name[:90]}")
