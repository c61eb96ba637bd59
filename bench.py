#!/usr/bin/env python
"""Headline benchmark: batched greedy caption decode on the nano configuration (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

One "step" = one `generate()` call: 8 captions (8 synthetic 224x224 images), prompt [[50256]], 64 new tokens, top_k=1,
KV cache, on-device no-repeat-n-gram ban.  Prints ONE JSON line (see DESIGN.md "measurement").
  value : tokens/s with the images already resident in HBM and the ids left on the device;
  e2e   : tokens/s through the public API with HOST buffers: pinned-host images copied to the device and the ids read
          back every step inside the timed region;
  roofline: the decode step against the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes per step stated in
          DESIGN.md; the dominant kernel (skinny weight-streaming linear) is also timed alone;
  cpu_baseline: the reference's own generate loop on this box's host cores, ONE step of the same workload.  It runs the
          UNMODIFIED reference (baseline/_ref, mirrored from /root/reference by __graft_entry__.build(); "kind":
          "reference") and falls back to the oracle port (oracle/i2t_oracle.py, "kind": "port") only when no copy of the
          reference is reachable;
  gpu_eager_reference: the unmodified reference in torch eager on the SAME B200 (fp32 with TF32 off, and bf16 autocast):
          one 8 x 64 generate and one B = 8 training step -- the bar SURVEY.md 8(d) names;
  train / roofline.secondary: the other half of BASELINE.json's metric (train img/s, nano + gpt2hf + momentum distillation).
`--impl reference` times the reference's CPU implementation on the same workload and config with all host threads.
With N > 1 (torchrun) every rank decodes its own 8 captions: independent units, no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAPTIONS, NEW_TOKENS, PROMPT = 8, 64, 50256


def read_peaks_full():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def read_peaks():
    d = read_peaks_full()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(self.rows)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def algorithmic_bytes_per_step(spec, batch, dtype_bytes, mean_len, greedy_fused=False):
    """Bytes one decode step MUST move: every decoder weight once, the self KV cache prefix, the cached
    cross K/V, the logits row, and the K/V append."""
    C, L, V, S = spec["n_embd"], spec["n_layer"], spec["vocab_size"], spec["n_cls"]
    ff = int(spec["ff_mult"] * C)
    n_cross = sum(1 for d in range(L) if d % 2 == 0 or not spec["skip_alternate_cross_attn"]) if spec["use_cross_attn"] else 0
    w = L * (3 * C * C + C * C + 2 * C * ff) + n_cross * (2 * C * C) + V * C          # GEMM weights read per step
    small = L * (3 * C + C + ff + C + 4 * C) * 4 + n_cross * (2 * C + 2 * C) * 4 + 2 * C * 4   # biases + LN params (fp32)
    kv_read = L * 2 * mean_len * C * batch * dtype_bytes
    kv_write = L * 2 * C * batch * dtype_bytes
    xkv = n_cross * 2 * S * C * batch * dtype_bytes
    # logits: written by the LM head and read by the sampler -- not moved at all when the arg-max is fused into the LM head
    logits = 0 if greedy_fused else batch * V * 4 * 2
    return w * dtype_bytes + small + kv_read + kv_write + xkv + logits, w * dtype_bytes


# FLOPs per image.  "model": every GEMM / attention of forward + backward as the reference executes them.
# "executed": what this implementation runs -- nano.yaml optimises 61 of 432 tensors; the weight gradients of the
# never-stepped parameters are skipped (outputs identical): ViT forward 35.1 + decoder forward 69.4 + data gradients 72.6 +
# weight gradients of the cross-attention / ln_3 / wpe / LSH parameters 3.0.  gpt2hf and the EMA teacher forward run in full.
TRAIN_CONFIGS = (
    # key, yaml, per-GPU batch, momentum distillation, model GFLOP / img, executed GFLOP / img
    ("nano_b8", "nano", 8, False, 243.4, 180.1),
    ("nano_b64", "nano", 64, False, 243.4, 180.1),
    ("nano_moco_b64", "nano", 64, True, 243.4 + 104.5, 180.1 + 104.5),
    ("gpt2hf_b32", "gpt2", 32, False, 340.0, 340.0),
)


def train_secondary(rank, world, local, steps=3, warmup=3, only=None):
    """The other half of BASELINE.json's metric (train img/s), measured live next to the headline on the same N GPUs: bf16, the
    YAML's dropout 0.1 and AdamW parameter groups, the YAML's gradient_accumulation_steps, CUDA-graphed micro-steps, data
    parallel over the ranks of this run (bucketed all-reduce behind in-graph events).  Configurations: nano.yaml at 8 (the YAML
    batch) and 64 images per GPU, nano.yaml with momentum distillation (BASELINE config 4), local/gpt2.yaml -- ViT-B/16 + HF
    GPT-2 layout, every weight trained (config 3).  W >= 3 warm-up steps, CUDA events, max over ranks."""
    import fnmatch
    import types
    from image2text_b200 import load_training_config
    from image2text_b200.config_schema import TrainerWrapperConfig
    from image2text_b200.dp import GradientAllReducer
    from image2text_b200.model_spec import synth_state_dict
    from image2text_b200.optimizer import AdamW
    from image2text_b200.synthetic import synth_images, synth_labels
    from image2text_b200.wrapper import ModelTrainerWrapper
    peaks = read_peaks_full()
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {"metric": "train img/s (bf16, dropout 0.1, AdamW on the YAML's parameter groups, the YAML's accumulation, CUDA-graphed "
                     "micro-steps, synthetic 224x224 images + random captions, data parallel over the ranks)", "unit": "img/s",
           "tensor_peak_tflops": sustained, "tensor_peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"
           if "bf16_tflops_sustained" in peaks else "fallback"}
    for key, yaml_name, bs, moco, gflop_model, gflop_exec in TRAIN_CONFIGS:
        if only is not None and key not in only:
            continue
        try:
            tc = load_training_config(os.path.join(ROOT, "configs", yaml_name + ".yaml"))
            accum = tc.gradient_accumulation_steps
            V = 50257
            tok = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, mask_token_id=None, vocab_size=V)
            tkw = dict(moco_momentum=0.995, moco_alpha=0.4) if moco else {}
            w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(**tkw), -100, device=f"cuda:{local}",
                                    compute_dtype=torch.bfloat16)
            w.model.load_state_dict(synth_state_dict(w.model.spec, seed=0))
            w.copy_momentum_params()
            w.model.set_dropout_seed(1234 + rank)
            w.train()
            groups, chosen = [], set()
            for oc in tc.optimizers:          # reference trainer.py:145-172
                ps = [p for n, p in w.named_parameters() if n.split(".", 1)[0] != "model_m" and
                      (oc.target_modules is None or any(fnmatch.fnmatch(n.split(".", 1)[-1], pat) for pat in oc.target_modules))]
                groups.append(dict(params=ps, lr=oc.lr, weight_decay=oc.weight_decay, betas=oc.betas))
                chosen.update(id(p) for p in ps)
            for _, p in w.model.named_parameters():
                if id(p) not in chosen:
                    p.requires_grad_(False)   # never stepped by the reference either: skip their weight gradients
            opt = AdamW(groups)
            red = GradientAllReducer([p for g in groups for p in g["params"]])
            red.attach_optimizer(opt)
            red.broadcast_parameters(w.model)
            w.copy_momentum_params()
            images = synth_images(bs, 224, seed=1234 + rank).cuda()
            labels = synth_labels(bs, 256, seed=1234 + rank).cuda()

            def one_step():
                for micro in range(accum):
                    with red.no_sync():
                        loss = w.train_step_graphed(images, labels, 1.0 / accum, reducer=red, sync=micro == accum - 1)
                red.finish()
                opt.step()
                opt.zero_grad(set_to_none=False)
                return loss

            losses = []
            for _ in range(max(warmup, 3)):
                losses.append(float(one_step()))
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = one_step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            imgs = bs * accum * steps * world
            ips = imgs / (ms / 1e3)
            out[key] = {"value": round(ips, 1), "ms_per_step": round(ms / steps, 2), "batch_per_gpu": bs, "accumulation": accum,
                        "gflop_per_img_model": gflop_model, "gflop_per_img_executed": gflop_exec,
                        "model_tflops": round(ips * gflop_model / 1e3, 1), "executed_tflops": round(ips * gflop_exec / 1e3, 1),
                        "executed_frac_of_tensor_peak": round(ips * gflop_exec / 1e3 / (sustained * world), 4),
                        "loss_first": round(losses[0], 4), "loss": round(float(loss), 4)}
            red.remove()
            del w, opt, red, images, labels
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    from image2text_b200 import VisionEncoderDecoder, load_training_config
    from image2text_b200._lib import call, launch_count
    from image2text_b200.model_spec import synth_state_dict
    from image2text_b200.synthetic import synth_images
    from image2text_b200 import ops

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
    model = VisionEncoderDecoder(tc.model, device=dev, compute_dtype=cd)
    model.load_state_dict(synth_state_dict(model.spec, seed=0))
    model.eval()
    spec = model.spec
    images_host = synth_images(CAPTIONS, 224, seed=1234 + rank).pin_memory()
    prompt_host = torch.full((CAPTIONS, 1), PROMPT, dtype=torch.long).pin_memory()
    images = images_host.to(dev)
    prompt = prompt_host.to(dev)
    out_host = torch.empty((CAPTIONS, 1 + NEW_TOKENS), dtype=torch.long).pin_memory()

    def step_resident():
        return model.generate(images, prompt, max_new_tokens=NEW_TOKENS, temperature=1.0, top_k=1, seed=0)

    def step_e2e():
        im = images_host.to(dev, non_blocking=True)
        pr = prompt_host.to(dev, non_blocking=True)
        ids = model.generate(im, pr, max_new_tokens=NEW_TOKENS, temperature=1.0, top_k=1, seed=0)
        out_host.copy_(ids, non_blocking=True)
        return ids

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local)
    sampler.start()
    eng = next(iter(model._decode_engines.values()))
    n0, g0 = launch_count(), eng.graph_launches
    ms = timed(step_resident, args.steps)
    eager_launches = launch_count() - n0 + (eng.graph_launches - g0)
    # kernels launched through the library directly (encoder, cross-KV prefill) + kernels inside the replayed step graphs
    launches = eager_launches + args.steps * eng.replays_last * (eng.launches_per_step or 0)
    ms_e2e = timed(step_e2e, args.steps)
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if eng.mode in ("mega2", "mega3"):
        # the decode kernel alone: ONE launch = the whole 64-token loop (cache length grows 1..64 inside it); mega3: the
        # poison fills of the exchange buffers / new cache rows (~25 MB of memset per launch) are inside the timed region
        reps = 10
        eng.generate(images, prompt, NEW_TOKENS, 1.0, 1, seed=0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            eng.pos.zero_()
            if eng.mode == "mega3":
                eng._mega3_run(0, NEW_TOKENS, 1.0, 1, 1)
            else:
                eng._mega2_run(0, NEW_TOKENS, 1.0, 1)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / (reps * NEW_TOKENS)
        mean_len = (NEW_TOKENS + 1) / 2
        kernel_desc = "decode_%s_kernel: one cooperative launch = %d decode steps; per-step figures" % (eng.mode, NEW_TOKENS)
    else:
        # decode step alone: replay the captured step graph (positions keep advancing inside the cache window)
        g = next(iter(eng.graphs.values()))
        eng.pos.fill_(8)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / reps
        mean_len = 8 + reps / 2
        kernel_desc = "decode step (one CUDA-graph replay, %d launches)" % eng.launches_per_step
    # the largest single weight stream alone (LM-head skinny linear as a separate kernel), back-to-back launches
    W = model.weights()
    wd = ops.F32 if cd == torch.float32 else ops.BF16
    V, C = spec["vocab_size"], spec["n_embd"]
    lm_w = W.c("decoder.lm_head.weight")
    st = ops.stream()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(40):
        call("i2t_dec_linear", ops.ptr(eng.x), ops.ptr(W["decoder.transformer.ln_f.weight"]),
             ops.ptr(W["decoder.transformer.ln_f.bias"]), 1e-5, ops.ptr(lm_w), None, None, ops.ptr(eng.logits), V, CAPTIONS, V, C, 0,
             wd, 0, None, None, 0, 0, 0, None, st)
    e1.record()
    torch.cuda.synchronize()
    lm_ms = e0.elapsed_time(e1) / 40
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    tokens = CAPTIONS * NEW_TOKENS * world
    value = tokens * args.steps / (ms / 1e3)
    e2e = tokens * args.steps / (ms_e2e / 1e3)
    esz = 2 if cd == torch.bfloat16 else 4
    step_bytes, weight_bytes = algorithmic_bytes_per_step(spec, CAPTIONS, esz, mean_len=mean_len, greedy_fused=eng.mode in ("mega2", "mega3"))
    peak, peak_src = read_peaks()
    achieved = step_bytes / (step_ms / 1e3) / 1e9
    lm_bytes = V * C * esz + CAPTIONS * V * 4
    lm = {"name": "dec_linear_kernel (LM head 50257x768 as a stand-alone launch; not on the mega2 path)",
          "algorithmic_bytes": int(lm_bytes), "us": round(lm_ms * 1e3, 2), "achieved": round(lm_bytes / (lm_ms / 1e3) / 1e9, 1),
          "frac": round(lm_bytes / (lm_ms / 1e3) / 1e9 / peak, 4)}
    if eng.mode == "mega3":
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel for this workload (one launch = 64 steps) from the
                    # committed ncu --set full capture
                    "traffic": 16800550608 if (cd == torch.bfloat16 and NEW_TOKENS == 64 and CAPTIONS == 8) else None,
                    "traffic_source": "ncu --set full, profiles/r02_ncu_full_decode_mega3_bf16_raw.csv (per launch of 64 steps)",
                    "peak_source": peak_src, "kernel": kernel_desc,
                    "algorithmic_bytes_per_launch": int(step_bytes * NEW_TOKENS), "us_per_launch": round(step_ms * 1e3 * NEW_TOKENS, 1),
                    "algorithmic_bytes_per_step": int(step_bytes), "us_per_step": round(step_ms * 1e3, 2),
                    "note": "dataflow megakernel: no grid barriers, poison-tagged exchange buffers, packed per-CTA weight streams "
                            "prefetched through a shared-memory ring (DESIGN.md section 8)"}
    elif eng.mode == "mega2":
        # the dominant (only) decode kernel: one launch = NEW_TOKENS steps.  traffic: dram__bytes_read.sum + dram__bytes_write.sum
        # of this kernel for this workload from the committed ncu --set full capture (profiles/r01_ncu_full_decode_mega2_bf16_raw.csv)
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": 17455036456 if (cd == torch.bfloat16 and NEW_TOKENS == 64 and CAPTIONS == 8) else None,
                    "traffic_source": "ncu --set full, profiles/r01_ncu_full_decode_mega2_bf16_raw.csv (per launch of 64 steps)",
                    "peak_source": peak_src, "kernel": kernel_desc,
                    "algorithmic_bytes_per_launch": int(step_bytes * NEW_TOKENS), "us_per_launch": round(step_ms * 1e3 * NEW_TOKENS, 1),
                    "algorithmic_bytes_per_step": int(step_bytes), "us_per_step": round(step_ms * 1e3, 2),
                    "note": "81 dependent stages per step separated by grid barriers (1.2 us floor each, scripts/micro/grid_barrier.cu): "
                            "latency-bound, see DESIGN.md section 8; LM-head stage alone: 77 MB in ~15 us (profiles/r01_trace_mega2_bf16.txt)"}
    else:
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": None, "peak_source": peak_src, "kernel": kernel_desc, "algorithmic_bytes_per_launch": int(step_bytes),
                    "us_per_launch": round(step_ms * 1e3, 2), "dominant_kernel": lm}
    line = {
        "metric": "decode tok/s (nano.yaml, 8 captions x 64 new tokens, greedy top_k=1, KV cache)",
        "value": round(value, 1), "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic (randn 224x224 images, seed 1234+rank; seeded random-init weights)",
        "config": {"workload": "training_configs/local/nano.yaml decode: 8 captions/step, 64 new tokens, prompt [[50256]], "
                               "top_k=1, no_repeat_n_grams [2,3,4,5]; one step = ViT-B/16 encode of 8 images + 64 decode steps",
                   "captions_per_gpu": CAPTIONS, "new_tokens": NEW_TOKENS,
                   "l2": "decoder weights per step (%.0f MB) exceed the 126 MB L2; no explicit flush" % (weight_bytes / 1e6),
                   "parallelism": f"replicated model, captions sharded by image over {world} GPU(s), no collective"},
        "e2e": {"value": round(e2e, 1), "unit": "tok/s", "h2d_bytes_per_step": int(images_host.numel() * 4 + prompt_host.numel() * 8),
                "d2h_bytes_per_step": int(out_host.numel() * 8), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": sampler.summary(),
    }
    del model
    torch.cuda.empty_cache()
    if not args.no_train:
        # every rank takes part (data parallel); a failure here must not cost the headline line
        try:
            line["train"] = train_secondary(rank, world, local)
        except Exception as e:  # noqa: BLE001
            line["train"] = {"error": repr(e)[:300]}
        # the driver keeps `roofline` and `config`: the tensor-bound half of BASELINE.json's metric rides there too
        keep = {k: {kk: v[kk] for kk in ("value", "ms_per_step", "executed_tflops", "executed_frac_of_tensor_peak", "model_tflops")
                    if kk in v} for k, v in line["train"].items() if isinstance(v, dict)}
        line["roofline"]["secondary"] = {"bound": "tensor", "unit": "TFLOP/s (executed, bf16)", "peak": line["train"].get("tensor_peak_tflops"),
                                         "what": "train img/s on the same GPUs, see `train`", **keep}
        line["config"]["train"] = {k: v.get("value") for k, v in keep.items()}
    if rank == 0 and world == 1 and not args.no_eager_ref:
        try:
            line["gpu_eager_reference"] = gpu_eager_reference(dev)
        except Exception as e:  # noqa: BLE001
            line["gpu_eager_reference"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sample_tokens=NEW_TOKENS)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _host_threads():
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time a single-threaded CPU
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    return torch.get_num_threads()


def load_reference_harness():
    """tests/golden/ref_harness.py if a copy of the UNMODIFIED reference is reachable (baseline/_ref, mirrored from
    /root/reference by __graft_entry__.build()), else None."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    try:
        import ref_harness
    except Exception:  # noqa: BLE001
        return None
    return ref_harness if ref_harness.find_reference_root() is not None else None


def reference_nano_model(rh, device="cpu"):
    """The unmodified reference VisionEncoderDecoder for nano.yaml with the SAME seeded weights as the B200 arm (loaded through
    the reference's own strict load_state_dict)."""
    import yaml
    from image2text_b200 import load_training_config
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
    spec = spec_from_config(tc.model)
    with open(os.path.join(ROOT, "configs", "nano.yaml")) as fh:
        cfg = yaml.safe_load(fh)
    model = rh.build_reference_model(cfg["model"], state_dict=synth_state_dict(spec, seed=0))
    return model.to(device).eval(), cfg


def cpu_generate_tok_s(new_tokens: int, steps: int = 1, warmup: int = 0, budget_s: float = None):
    """The reference's cache-less generate loop on the host cores: 8 captions x `new_tokens`, greedy.  Returns
    (tok/s, seconds per step, kind, steps run).  `budget_s` bounds the timed region (steps are cut, never the workload)."""
    from image2text_b200 import load_training_config
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    from image2text_b200.synthetic import synth_images
    _host_threads()
    images = synth_images(CAPTIONS, 224, seed=1234)
    prompt = torch.full((CAPTIONS, 1), PROMPT, dtype=torch.long)
    rh = load_reference_harness()
    if rh is not None:
        model, _ = reference_nano_model(rh, "cpu")
        kind = "reference"

        def one():
            return model.generate(images, prompt, max_new_tokens=new_tokens, temperature=1.0, top_k=1)
    else:
        from oracle import i2t_oracle as O
        tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
        spec = spec_from_config(tc.model)
        sd = synth_state_dict(spec, seed=0)
        kind = "port"

        def one():
            return O.generate(sd, spec, images, prompt, new_tokens, top_k=1)
    for _ in range(warmup):
        one()
    done, t0 = 0, time.perf_counter()
    while done < steps:
        one()
        done += 1
        if budget_s is not None and done < steps and (time.perf_counter() - t0) * (done + 1) / done > budget_s:
            break
    dt = time.perf_counter() - t0
    return CAPTIONS * new_tokens * done / dt, dt / done, kind, done


def cpu_baseline(sample_tokens: int):
    v, sec, kind, _ = cpu_generate_tok_s(sample_tokens)
    what = ("the UNMODIFIED reference (baseline/_ref) VisionEncoderDecoder.generate" if kind == "reference"
            else "oracle/i2t_oracle.generate (port of the reference's cache-less loop; no copy of the reference reachable)")
    return {"value": round(v, 2), "unit": "tok/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{what}, ONE step of the same workload: 8 captions x {sample_tokens} new tokens incl. ViT encode, fp32, "
                      f"{sec:.1f} s", "host_cpus": os.cpu_count()}


def gpu_eager_reference(dev):
    """SURVEY.md 8(d) "the real bar": the UNMODIFIED reference in torch eager on this B200 -- fp32 (TF32 off, torch's default)
    and bf16 autocast (what accelerate's mixed_precision='bf16' does) -- for one 8 x 64 greedy generate and one nano.yaml
    training step of 8 images (train_step + backward + AdamW on the YAML's parameter groups, dropout 0.1)."""
    import fnmatch
    import types
    from image2text_b200.synthetic import synth_images, synth_labels
    rh = load_reference_harness()
    if rh is None:
        return {"unavailable": "no copy of the reference reachable (baseline/_ref absent)"}
    out = {"impl": "unmodified reference, torch %s eager, same seeded weights / inputs" % torch.__version__}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, cfg = reference_nano_model(rh, dev)
    images = synth_images(CAPTIONS, 224, seed=1234).to(dev)
    prompt = torch.full((CAPTIONS, 1), PROMPT, dtype=torch.long, device=dev)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        def gen():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                return model.generate(images, prompt, max_new_tokens=NEW_TOKENS, temperature=1.0, top_k=1)
        ms = timed(gen, 2)
        out["generate_8x64_" + name] = {"value": round(CAPTIONS * NEW_TOKENS / (ms / 1e3), 1), "unit": "tok/s", "ms": round(ms, 1)}
    del model
    ref = rh.load_reference()
    tcfg = ref.CM.VisionEncoderDecoderConfig.model_validate(_offline_model_cfg(cfg["model"]))
    tok = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, mask_token_id=None, vocab_size=50257)
    from image2text_b200 import load_training_config
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    tc = load_training_config(os.path.join(ROOT, "configs", "nano.yaml"))
    w = ref.TW.ModelTrainerWrapper(tcfg, tok, ref.CT.TrainerWrapperConfig(), -100)
    w.model.load_state_dict(synth_state_dict(spec_from_config(tc.model), seed=0), strict=True)
    w = w.to(dev).train()
    groups = []
    for oc in tc.optimizers:          # reference trainer.py:145-172
        ps = [p for n, p in w.named_parameters() if n.split(".", 1)[0] != "model_m" and
              (oc.target_modules is None or any(fnmatch.fnmatch(n.split(".", 1)[-1], pat) for pat in oc.target_modules))]
        groups.append(dict(params=ps, lr=oc.lr, weight_decay=oc.weight_decay, betas=oc.betas))
    opt = torch.optim.AdamW(groups)
    bs, accum = tc.batch_size, tc.gradient_accumulation_steps
    timg = synth_images(bs, 224, seed=1234).to(dev)
    tlab = synth_labels(bs, 256, seed=1234).to(dev)
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        def step():
            for _ in range(accum):                         # training/utils.py:85-101 (accelerate-free restatement)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    loss, _ = w.train_step(timg, tlab)
                (loss / accum).backward()
            opt.step()
            opt.zero_grad()
        ms = timed(step, 2)
        out["train_nano_b8_" + name] = {"value": round(bs * accum / (ms / 1e3), 1), "unit": "img/s", "ms_per_step": round(ms, 1),
                                        "accumulation": accum}
    return out


def _offline_model_cfg(model_cfg):
    import copy
    d = copy.deepcopy(model_cfg)
    dec = d["decoder_config"]
    if "pretrained_model" in dec:
        dec["pretrained_model"] = None
    if "lora_spec" in dec:
        dec["lora_spec"] = None
    if "lora_spec" in d["vision_encoder_config"]:
        d["vision_encoder_config"]["lora_spec"] = None
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the SAME workload and config as the B200 arm (8 captions x 64 new tokens per step); a step costs 10-30 s of host time, so
    # the timed region is bounded by cutting STEPS (never the workload): at most ~150 s, at least one step, one warm-up
    warm = min(args.warmup, 1)
    v, sec, kind, done = cpu_generate_tok_s(NEW_TOKENS, steps=args.steps, warmup=warm, budget_s=150.0)
    what = ("UNMODIFIED reference (baseline/_ref) VisionEncoderDecoder.generate" if kind == "reference"
            else "oracle port of the reference's generate loop (no copy of the reference reachable)")
    line = {
        "impl": "reference", "metric": "decode tok/s (nano.yaml, 8 captions x 64 new tokens, greedy top_k=1, KV cache)",
        "value": round(v, 2), "unit": "tok/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": done,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic (same images / weights as the B200 arm)",
        "config": {"workload": "training_configs/local/nano.yaml decode: 8 captions/step, 64 new tokens, prompt [[50256]], "
                               "top_k=1, no_repeat_n_grams [2,3,4,5]; one step = ViT-B/16 encode of 8 images + 64 decode steps",
                   "captions_per_gpu": CAPTIONS, "new_tokens": NEW_TOKENS,
                   "note": f"{what}: the reference has no KV cache (every step re-runs the decoder over the whole prefix); host "
                           f"CPU, all threads; timed steps cut to {done} of {args.steps} to bound the run"},
        "cpu_baseline": {"value": round(v, 2), "unit": "tok/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{done} step(s) of 8 captions x {NEW_TOKENS} new tokens, {what}, fp32"},
        "e2e": {"value": round(v, 2), "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("I2T_BENCH_DTYPE", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary train img/s measurement")
    ap.add_argument("--no-eager-ref", action="store_true", help="skip the unmodified reference in torch eager on the GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
