"""Per-launch timeline (CUPTI) of the encode part of generate() for the bench workload (8 images, nano, bf16): ViT trunk + LSH tail
+ cross K/V prefill, i.e. everything before the decode kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=torch.bfloat16)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
images = synth_images(8, 224, seed=1234).cuda()
prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(3):
    m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=1, seed=1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.generate(images, prompt, max_new_tokens=64, temperature=1.0, top_k=1, seed=1)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
idx = next(i for i, e in enumerate(evs) if "decode_mega2" in e.name)
print(f"{idx} launches before the decode kernel, {evs[idx].time_range.start - t0:.1f} us; decode kernel {evs[idx].time_range.end - evs[idx].time_range.start:.1f} us")
agg = {}
for e in evs[:idx]:
    k = e.name[:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t:9.1f} us  x{n:4d}  {k}")
n_show = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for e in evs[:n_show]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:7.2f} us  {e.name[:80]}")
