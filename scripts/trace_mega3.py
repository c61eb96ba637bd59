"""Per-stage timing of the dataflow decode megakernel (clock64 stamps of chosen CTAs for the LAST step):
    python scripts/trace_mega3.py [cta ...]
For every stage of the step a CTA records: entry, inputs staged (linear stages it takes part in), done.  A stage a CTA
does not take part in costs it nothing (no barrier), so `wait` = entry -> staged is the time spent polling for inputs +
LayerNorm, `work` = staged -> done is MMAs + reduction + epilogue."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import VisionEncoderDecoder, load_training_config  # noqa: E402
from image2text_b200.decode_engine import DecodeEngine  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images  # noqa: E402

ctas = [int(a) for a in sys.argv[1:]] or [0, 37, 101, 147]
tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
m = VisionEncoderDecoder(tc.model, device="cuda", compute_dtype=torch.bfloat16)
m.load_state_dict(synth_state_dict(m.spec, seed=0))
m.eval()
eng = DecodeEngine(m, 8, mode="mega3")
images = synth_images(8, 224, seed=1234).cuda()
prompt = torch.full((8, 1), 50256, dtype=torch.long, device="cuda")
for _ in range(3):
    eng.generate(images, prompt, 64, 1.0, 1, seed=0)
T = eng._mega3
print("packed weight streams: %.1f MB over %d CTAs, exchange buffers %d KB x 3 generations" %
      (T["packed_bytes"] / 1e6, T["grid"], T["gen_stride"] // 1024))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    eng.pos.zero_()
    e0.record()
    eng._mega3_run(0, 64, 1.0, 1, 1)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / 64)
print("decode loop, CUDA events: %.1f us per step (best of 5; all: %s)" % (min(ts), " ".join("%.1f" % t for t in ts)))

def loop_us(n):
    best = 1e30
    for _ in range(3):
        eng.pos.zero_()
        e0.record()
        eng._mega3_run(0, n, 1.0, 1, 1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


t64, t128 = loop_us(64), loop_us(128)
print("one launch of 64 steps %.0f us, of 128 steps %.0f us -> %.1f us per additional step, %.0f us of per-launch overhead "
      "(poison fills, cooperative launch, table copy, first-step effects)" % (t64, t128, (t128 - t64) / 64, t64 - (t128 - t64)))

sched = T["sched"].cpu()
lin = T["lin"].cpu()
n_sched = sched.shape[0]
mhz = 1965.0
for cta in ctas:
    os.environ["I2T_TRACE_CTA"] = str(cta)
    eng.trace = torch.zeros(n_sched * 8, dtype=torch.int64, device="cuda")
    eng.generate(images, prompt, 32, 1.0, 1, seed=0)      # stamps of the LAST step (cache length 32)
    torch.cuda.synchronize()
    tr = eng.trace.view(n_sched, 8).cpu()
    t0 = int(tr[0, 0])
    tot = {}
    rows = []
    for s in range(n_sched):
        kind, idx = int(sched[s, 0]), int(sched[s, 1])
        b, st, dn = int(tr[s, 0]), int(tr[s, 1]), int(tr[s, 2])
        if kind == 0:
            name = f"lin N={int(lin[idx, 7])} K={int(lin[idx, 8])}"
        elif kind == 1:
            name = "attn"
        else:
            continue
        took = st > 0 if kind == 0 else (dn - b) > 400
        wait = (st - b) / mhz if (kind == 0 and st > 0) else 0.0
        work = (dn - st) / mhz if (kind == 0 and st > 0) else (dn - b) / mhz
        a = tot.setdefault(name, [0, 0, 0.0, 0.0])
        a[0] += 1
        if took:
            a[1] += 1
            a[2] += wait
            a[3] += work
        fine = ""
        if kind == 0 and st > 0:
            polled, spins, wr, mm, cs = int(tr[s, 6]), int(tr[s, 7]), int(tr[s, 3]), int(tr[s, 4]), int(tr[s, 5])
            if int(lin[idx, 17]) & 1:
                print("  LM head: %.2f us of the stage parked on the weight ring after the first tile" % (spins / mhz))
            fine = "  poll %5.2f (%4d spins) ln+stage %5.2f | weights %5.2f mma %5.2f sync %5.2f epilogue+rest %5.2f" % (
                ((polled - b) / mhz if polled > 0 else 0.0), spins, ((st - polled) / mhz if polled > 0 else (st - b) / mhz),
                (wr - st) / mhz, (mm - wr) / mhz, (cs - mm) / mhz, (dn - cs) / mhz)
        rows.append((s, name, took, (b - t0) / mhz, wait, work, fine))
    print("  step begin -> tokens known: %.2f us (waiting for every CTA's LM-head keys of the previous step)" %
          ((int(tr[n_sched - 1, 4]) - int(tr[n_sched - 1, 3])) / mhz))
    last = int(tr[n_sched - 2, 2])
    print(f"--- CTA {cta}: step (first stage entry -> LM head done) {(last - t0) / mhz:.1f} us")
    for s, name, took, at, wait, work, fine in rows[:20] + rows[-3:]:
        print(f"  stage {s:3d} {name:22s} {'run ' if took else 'skip'} at {at:7.2f} us  wait {wait:6.2f}  work {work:6.2f}{fine}")
    for k, (n, nt, w, c) in tot.items():
        if nt:
            print(f"  {k:22s} x{n:3d} (took part in {nt:3d})  wait {w / nt:6.2f}  work {c / nt:6.2f}  sum {w + c:8.1f} us")
