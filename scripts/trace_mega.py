"""Per-stage timing of the decode megakernel (clock64 stamps of CTA 0): python scripts/trace_mega.py [fp32|bf16] [mega|mega2]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=dtype)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
mode = sys.argv[2] if len(sys.argv) > 2 else ("mega2" if dtype == torch.bfloat16 else "mega")
eng = DecodeEngine(m, 8, mode=mode)
images = synth_images(8, 224, seed=1234).cuda()
prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
eng.generate(images, prompt, 8, 1.0, 1, seed=0)
n_sched = eng._mega["sample"].shape[0]
W = 8 if mode == "mega2" else 4
eng.trace = torch.zeros(n_sched * W, dtype=torch.int64, device="cuda")
if mode == "mega2":
    # wall time of the decode loop alone (64 steps, one launch), CUDA events
    for _ in range(3):
        eng.generate(images, prompt, 64, 1.0, 1, seed=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        eng.pos.zero_()
        e0.record()
        eng._mega2_run(0, 64, 1.0, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 64)
    print("decode loop, CUDA events: %.1f us per step (best of 5; all: %s); SM clock now %s MHz" %
          (min(ts), " ".join("%.1f" % t for t in ts), torch.cuda.clock_rate()))
    eng.generate(images, prompt, 32, 1.0, 1, seed=0)      # stamps of the LAST step (cache length 32)
else:
    for _ in range(3):
        eng._mega_step(True, 1.0, 1)
torch.cuda.synchronize()
tr = eng.trace.view(n_sched, W).cpu()
fine = {}
sched = eng._mega["sample"].cpu()
lin = eng._mega["lin"].cpu()
mhz = 1965.0
t0 = int(tr[0, 0])
tot = {}
print("stage kind idx   N     K   | stage_x  compute  barrier  total (us)")
for s in range(n_sched):
    kind, idx = int(sched[s, 0]), int(sched[s, 1])
    b, st, cp, sy = [int(v) for v in tr[s][:4]]
    if min(b, st, cp, sy) <= 0 or not (b <= st <= cp <= sy):
        continue            # CTA 0 owns no unit of this stage: some of its clock64 stamps were never written
    if W == 8 and kind == 0 and int(tr[s][1]) > b:
        N, K = int(lin[idx, 7]), int(lin[idx, 8])
        ld, arr, pf, mma = [int(v) for v in tr[s][4:8]]
        f = fine.setdefault(f"lin N={N} K={K}", [0, 0., 0., 0., 0., 0., 0., 0.])
        f[0] += 1
        for i, v in enumerate([(ld - b) if ld > b else 0, (st - ld) if ld > b else (st - b), mma - st, cp - mma, arr - cp, pf - arr, sy - pf]):
            f[i + 1] += v / mhz
    if kind == 0:
        N, K = int(lin[idx, 7]), int(lin[idx, 8])
        row = ((st - b) / mhz, (cp - st) / mhz, (sy - cp) / mhz)
        name = f"lin N={N} K={K}"
    elif kind == 1:
        N = K = 0
        row = (0.0, (cp - b) / mhz, (sy - cp) / mhz)
        name = "attn"
    else:
        continue
    a = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    for i in range(3):
        a[i + 1] += row[i]
    if s < 12 or s > n_sched - 4:
        print(f"{s:4d} {kind:4d} {idx:4d} {N:5d} {K:5d} | {row[0]:7.2f} {row[1]:8.2f} {row[2]:8.2f} {sum(row):7.2f}")
last = int(tr[n_sched - 2, 3])
print("whole step (CTA 0, to the last barrier): %.1f us" % ((last - t0) / mhz))
for k, (n, a, b, c) in tot.items():
    print(f"{k:20s} x{n:3d}  stage_x {a / n:6.2f}  compute {b / n:6.2f}  barrier {c / n:6.2f}  sum {(a + b + c):8.1f} us")

if fine:
    print("fine (CTA 0 has a unit): x loads+LN params | LN+pack+sync+xf | first MMA | reduce+epilogue | arrive | prefetch issue | wait")
    for k, f in fine.items():
        print(f"{k:20s} x{f[0]:3d} " + " ".join(f"{v / f[0]:7.2f}" for v in f[1:]))
