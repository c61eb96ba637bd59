"""Short decode used under ncu: nano config, 8 captions, a few new tokens.  python scripts/profile_decode.py [dtype] [tokens]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=dtype)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
images = synth_images(8, 224, seed=1234).cuda()
prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
out = m.generate(images, prompt, max_new_tokens=tokens, top_k=1)
torch.cuda.synchronize()
print(out[:, :tokens + 1].tolist()[0])
