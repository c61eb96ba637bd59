"""Secondary benchmark (BASELINE.json configs[0]/[3] shape): nano.yaml training step on ONE or N GPUs.

    python scripts/bench_train.py [--dtype fp32|bf16] [--steps K] [--moco] [--cpu-ref]
    torchrun --nproc-per-node N scripts/bench_train.py ...

One optimiser-visible step = gradient_accumulation_steps (4) micro-steps of batch 8 per GPU: train_step -> backward ->
[DP all-reduce overlapped with backward] -> fused AdamW on the YAML's two parameter groups (reference trainer.py:145-172;
loop restated from training/utils.py:85-101).  Dropout runs at the YAML's values (0.1: transformer.drop, token-level q/k/v,
SDPA dropout_p, resid / MLP dropout, cross-attention dropout) unless --no-dropout is given.
Prints one JSON line: img/s (all ranks), ms/step, achieved model TFLOP/s (243.4 GFLOP/img).
"""
import argparse
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--moco", action="store_true")
    ap.add_argument("--no-dropout", action="store_true", help="switch every dropout off (the parity configuration)")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--profile-glue", action="store_true", help="list the torch (aten) ops of one eager micro-step by call site")
    ap.add_argument("--config", default="nano", choices=["nano", "gpt2"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU micro-batch (0 = the YAML value)")
    ap.add_argument("--graph", type=int, default=1, help="1: forward+backward of a micro-step replayed as a CUDA graph")
    ap.add_argument("--cpu-ref", action="store_true", help="time the CPU oracle's train step instead")
    args = ap.parse_args()
    from image2text_b200 import load_training_config
    from image2text_b200.config_schema import TrainerWrapperConfig
    from image2text_b200.model_spec import spec_from_config, synth_state_dict
    from image2text_b200.synthetic import synth_images, synth_labels
    tc = load_training_config(os.path.join(ROOT, "configs", args.config + ".yaml"))
    bs, accum = (args.batch or tc.batch_size), tc.gradient_accumulation_steps
    if args.cpu_ref:
        from oracle import i2t_oracle as O
        spec = spec_from_config(tc.model)
        sd = synth_state_dict(spec, seed=0)
        train_keys = [k for k in sd if ("lsh_emb" in k and k.endswith("emb.weight")) or "wpe" in k or "cross_attn" in k or "ln_3" in k]
        sd = {k: (v.clone().requires_grad_(True) if k in train_keys else v) for k, v in sd.items()}
        images, labels = synth_images(bs, 224, seed=1234), synth_labels(bs, 256, seed=1234)
        t0 = time.perf_counter()
        loss = O.train_step_loss(sd, spec, images, labels)
        loss.backward()
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "cpu oracle (port of the reference train step, dropout 0)", "img_per_s": round(bs / dt, 3),
                          "s_per_micro_step": round(dt, 2), "threads": torch.get_num_threads(), "loss": float(loss)}))
        return
    from image2text_b200.dp import GradientAllReducer
    from image2text_b200.optimizer import AdamW
    from image2text_b200.wrapper import ModelTrainerWrapper
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tok = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, mask_token_id=None, vocab_size=50257)
    tkw = dict(moco_momentum=0.995, moco_alpha=0.4) if args.moco else {}
    over = dict(dropout=0.0, attn_dropout=0.0) if args.no_dropout else {}
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(**tkw), -100, device=f"cuda:{local}", compute_dtype=cd,
                            spec_overrides=over)
    w.model.set_dropout_seed(1234 + rank)
    w.model.load_state_dict(synth_state_dict(w.model.spec, seed=0))
    w.copy_momentum_params()
    w.train()
    # param groups exactly as reference trainer.py:145-172 (fnmatch on the name without the leading "model.")
    import fnmatch
    groups, chosen = [], set()
    for oc in tc.optimizers:
        ps = [p for n, p in w.named_parameters() if n.split(".", 1)[0] != "model_m" and
              (oc.target_modules is None or any(fnmatch.fnmatch(n.split(".", 1)[-1], pat) for pat in oc.target_modules))]
        groups.append(dict(params=ps, lr=oc.lr, weight_decay=oc.weight_decay, betas=oc.betas))
        chosen.update(id(p) for p in ps)
    # Q5: parameters outside every group never change; skipping their weight gradients is a pure saving
    for n, p in w.model.named_parameters():
        if id(p) not in chosen:
            p.requires_grad_(False)
    opt = AdamW(groups)
    # I2T_DP_OVERLAP=1: in-graph bucket events (exchange overlapped with the backward; measured slower, see dp.py)
    red = GradientAllReducer([p for g in groups for p in g["params"]], graph_events=os.environ.get("I2T_DP_OVERLAP", "0") == "1")
    red.attach_optimizer(opt)
    red.broadcast_parameters(w.model)
    w.copy_momentum_params()
    images = synth_images(bs, 224, seed=1234 + rank).cuda()
    labels = synth_labels(bs, 256, seed=1234 + rank).cuda()

    sections = os.environ.get("I2T_BENCH_SECTIONS") == "1"          # per-section device times of every step (rank 0 prints)
    sec_ev = []

    def one_step():
        if sections:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        for micro in range(accum):
            last = micro == accum - 1
            # graphed micro-steps never fire the all-reduce hooks (a replay runs no Python): with --graph every micro-step is
            # a replay and red.finish() exchanges all buckets after the last backward (NVLink moves the ~0.1-1 GB of fp32
            # gradients in a few ms; an eager last micro-step would cost 10x that in launch overhead).  --graph 0: eager
            # micro-steps, bucketed all-reduce overlapped with the last backward.
            ctx = red.no_sync() if (not last or args.graph) else torch.enable_grad()
            with ctx:
                if args.graph:
                    loss = w.train_step_graphed(images, labels, 1.0 / accum, reducer=red, sync=last)
                else:
                    loss, _ = w.train_step(images, labels)
                    (loss / accum).backward()
        if sections:
            ev[1].record()
        red.finish()
        if sections:
            ev[2].record()
        opt.step()
        opt.zero_grad(set_to_none=False)
        if sections:
            ev[3].record()
            sec_ev.append(ev)
        return loss

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if args.profile and world > 1:
        raise SystemExit("--profile runs one extra step on rank 0 only: single process only")
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one_step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
    if args.profile_glue and rank == 0:
        # every aten op of one eager micro-step (forward + backward) with the Python line that issued it; pure view ops are skipped
        import traceback
        from torch.utils._python_dispatch import TorchDispatchMode
        views = {"view", "_unsafe_view", "slice", "select", "detach", "alias", "as_strided", "t", "transpose", "permute", "expand",
                 "unsqueeze", "squeeze", "reshape", "empty", "empty_like", "empty_strided", "new_empty", "_local_scalar_dense",
                 "unbind", "split", "split_with_sizes", "narrow", "view_as", "is_same_size", "stride", "size", "numel"}
        sites = {}

        class Log(TorchDispatchMode):
            def __torch_dispatch__(self, func, types, a=(), kw=None):
                name = func.__name__.split(".")[0]
                if name not in views:
                    fr = [f for f in traceback.extract_stack() if "image2text_b200" in f.filename]
                    site = f"{os.path.basename(fr[-1].filename)}:{fr[-1].lineno} {fr[-1].name}" if fr else "autograd engine"
                    sites[(name, site)] = sites.get((name, site), 0) + 1
                return func(*a, **(kw or {}))

        from image2text_b200 import ops
        with Log(), ops.grad_sinks():                 # what a captured micro-step issues (wrapper.train_step_graphed)
            loss, _ = w.train_step(images, labels)
            (loss / accum).backward()
            torch.cuda.synchronize()
        for (name, site), n in sorted(sites.items(), key=lambda kv: -kv[1]):
            print(f"{n:5d} x {name:28s} {site}")
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if sections and rank == 0:
        e = sec_ev[-1]
        print(f"sections of the last step (ms, device time on the main stream): micro-steps {e[0].elapsed_time(e[1]):.2f}, "
              f"reducer.finish {e[1].elapsed_time(e[2]):.2f}, optimizer {e[2].elapsed_time(e[3]):.2f}")
    imgs = bs * accum * args.steps * world
    gflop_img = (243.4 + (104.5 if args.moco else 0.0)) if args.config == "nano" else (340.0 + (113.4 if args.moco else 0.0))
    if rank == 0:
        print(json.dumps({"metric": f"train img/s ({args.config}.yaml, B={bs}/GPU, accum {accum}, AdamW, dropout "
                                    f"{'off' if args.no_dropout else w.model.spec['dropout']})",
                          "value": round(imgs / (ms / 1e3), 2), "n_gpus": world, "dtype": args.dtype, "moco": args.moco, "graph": args.graph,
                          "ms_per_step": round(ms / args.steps, 2), "model_tflops": round(imgs * gflop_img / (ms / 1e3) / 1e3, 2),
                          "loss": float(loss), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
