"""One short generate() through the dataflow decode megakernel (for ncu): python scripts/profile_mega3.py [new_tokens]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=torch.bfloat16)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
eng = DecodeEngine(m, 8, mode=os.environ.get("I2T_DECODE", "mega3"))
images = synth_images(8, 224, seed=1234).cuda()
prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(2):
    out = eng.generate(images, prompt, n, 1.0, 1, seed=0)
torch.cuda.synchronize()
print("ok", out.shape, int(out.sum()))
