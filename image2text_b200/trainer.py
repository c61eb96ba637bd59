"""accelerate-free mirror of the reference's training entry point (``trainer.py:105-208``) and loops
(``training/utils.py:63-162``) over the B200 path, so that ``python -m image2text_b200.trainer --config_file <yaml>
--chkpt_file <pt>`` -- or ``torchrun --nproc-per-node N -m image2text_b200.trainer ...`` -- does what
``accelerate launch trainer.py`` does for the same YAML and the same checkpoint file:

* parameter groups from the YAML's ``optimizers[*].target_modules`` globs (``PatternMatcher`` = fnmatch on the parameter
  name without the leading ``model.``; ``model_m`` -- the EMA teacher -- is never optimised), ``SNRAdam`` when
  ``use_snr_optim`` else ``AdamW`` (reference trainer.py:145-172), here the fused one-launch-per-group steps;
* ``train_loop``: ``num_steps`` micro-batches per epoch, loss / accum backward, optimiser step + zero_grad every
  ``gradient_accumulation_steps`` micro-batches (what ``accelerator.accumulate`` does), MoCo reset epochs, checkpoint after
  every epoch holding ONLY the matched parameters when target globs are given (training/utils.py:104-123);
* ``val_loop`` (mean loss / metrics over ``num_val_steps``, all-reduced over ranks) and ``eval_model`` (one image, several
  candidates, temperature 0.7, nucleus 0.6: trainer.py:27-65);
* data parallelism is a real mean all-reduce of the gradients (the reference's DDP wrap never fires).

What is NOT here: the Flickr30K / deeplake dataloader and the hub tokenizer (no network; out of the hot-path scope).  The
data source is any iterator of ``(images, labels)`` batches; ``--synthetic`` (the only built-in one) yields batches of the
shapes ``trainer.py:68-103`` produces.  Token ids instead of decoded text are printed when no tokenizer is given.
"""
from __future__ import annotations

import argparse
import fnmatch
import os
import types
from typing import Iterator, List, Optional, Sequence, Tuple

import torch

from .config_schema import TrainingConfig, load_training_config
from .dp import GradientAllReducer
from .optimizer import AdamW, SNRAdam
from .synthetic import synth_images, synth_labels
from .wrapper import ModelTrainerWrapper


class PatternMatcher:
    """reference models/utils.py:16-28: any(fnmatch(candidate, pattern))."""

    def __init__(self, patterns: Sequence[str]):
        self.patterns = list(patterns)

    def match(self, candidate: str) -> bool:
        return any(fnmatch.fnmatch(candidate, p) for p in self.patterns)


def build_param_groups(model_wrapper: ModelTrainerWrapper, config: TrainingConfig):
    """reference trainer.py:145-167.  Returns (param_groups, matchers)."""
    groups, matchers = [], []
    for oc in config.optimizers:
        if oc.target_modules is not None:
            matcher = PatternMatcher(oc.target_modules)
            params = [p for n, p in model_wrapper.named_parameters()
                      if n.split(".", 1)[0] != "model_m" and matcher.match(n.split(".", 1)[-1])]
            matchers.append(matcher)
        else:
            assert len(config.optimizers) == 1
            params = [p for n, p in model_wrapper.named_parameters() if not n.startswith("model_m.")]
        groups.append({"lr": oc.lr, "weight_decay": oc.weight_decay, "betas": tuple(oc.betas), "params": params})
    return groups, matchers


def checkpoint_state(model: torch.nn.Module, matchers: List[PatternMatcher]):
    """What the reference writes after an epoch (training/utils.py:111-123): the matched PARAMETERS only when there are
    target globs (buffers never match: only named_parameters() are scanned), else the whole state dict."""
    sd = model.state_dict()
    if not matchers:
        return sd
    return {k: sd[k] for k, _ in model.named_parameters() if any(m.match(k) for m in matchers)}


def save_checkpoint(model: torch.nn.Module, path: str, matchers: List[PatternMatcher]):
    tmp = path + ".tmp"
    torch.save({k: v.detach().cpu() for k, v in checkpoint_state(model, matchers).items()}, tmp)
    os.replace(tmp, path)


def _is_main() -> bool:
    return int(os.environ.get("RANK", "0")) == 0


def _print(*a):
    if _is_main():
        print(*a, flush=True)


def train_loop(model_wrapper: ModelTrainerWrapper, optimizer, train_iter: Iterator[Tuple[torch.Tensor, torch.Tensor]],
               epoch: int, num_steps: Optional[int], accum: int = 1, reducer: Optional[GradientAllReducer] = None,
               reset_moco_after_k_epochs: Optional[List[int]] = None, chckpt_fname: Optional[str] = None,
               matchers: Sequence[PatternMatcher] = (), graph: bool = False, log_every: int = 0,
               state: Optional[dict] = None) -> bool:
    """reference training/utils.py:63-123.  `state["micro"]` carries the accumulation phase across epochs like accelerate's
    step counter does.  graph=True replays forward + backward of a micro-step as one CUDA graph (fixed shapes)."""
    model_wrapper.train()
    device = next(model_wrapper.model.parameters()).device
    state = state if state is not None else {}
    stop = False
    num_steps = 100 if num_steps is None else num_steps
    last = None
    loss_sum, n_done = None, 0
    for step in range(num_steps):
        try:
            images, labels = next(train_iter)
        except StopIteration:
            stop = True
            break
        images, labels = images.to(device, non_blocking=True), labels.to(device, non_blocking=True)
        micro = state.get("micro", 0) + 1
        sync = micro % accum == 0
        state["micro"] = micro
        ctx = reducer.no_sync() if (reducer is not None and (not sync or graph)) else torch.enable_grad()
        with ctx:
            if graph:
                loss = model_wrapper.train_step_graphed(images, labels, 1.0 / accum, reducer=reducer, sync=sync)
                metrics = {"train_loss_lm": loss}
            else:
                loss, metrics = model_wrapper.train_step(images, labels)
                (loss / accum).backward()
        if sync:
            if reducer is not None:
                reducer.finish()
            optimizer.step()
            optimizer.zero_grad(set_to_none=False)
        last = metrics
        lm = metrics["train_loss_lm"].detach().float()
        loss_sum = lm.clone() if loss_sum is None else loss_sum + lm      # on the device: no per-step synchronisation
        n_done += 1
        if log_every and (step + 1) % log_every == 0:
            _print(f"Epoch: {epoch} step {step + 1}/{num_steps} " + " ".join(f"{k}={float(v):.4f}" for k, v in metrics.items()))
    if last is not None:
        _print(f"Epoch: {epoch} " + " ".join(f"{k}={float(v):.4f}" for k, v in last.items()))
        state["last_train_loss"] = float(loss_sum) / n_done               # mean over the epoch
    if reset_moco_after_k_epochs is not None and (epoch + 1) in reset_moco_after_k_epochs:
        model_wrapper.copy_momentum_params()
    if chckpt_fname is not None:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()
        if _is_main():
            save_checkpoint(model_wrapper.model, chckpt_fname, list(matchers))
    return stop


@torch.no_grad()
def val_loop(model_wrapper: ModelTrainerWrapper, val_iter, epoch: int, num_val_steps: Optional[int]):
    """reference training/utils.py:126-162: mean loss / metrics over the steps (and over ranks)."""
    model_wrapper.eval()
    device = next(model_wrapper.model.parameters()).device
    num_steps = 100 if num_val_steps is None else num_val_steps
    losses, metrics_all = [], {}
    for _ in range(num_steps):
        images, labels = next(val_iter)
        loss, metrics = model_wrapper.val_step(images.to(device), labels.to(device))
        losses.append(loss.detach().float())
        for k, v in metrics.items():
            metrics_all[k] = metrics_all.get(k, 0.0) + v.detach().float() / num_steps
    loss = torch.stack(losses).mean()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        world = torch.distributed.get_world_size()
        pack = torch.stack([loss] + [metrics_all[k] for k in sorted(metrics_all)])
        torch.distributed.all_reduce(pack)
        pack /= world
        loss = pack[0]
        metrics_all = {k: pack[1 + i] for i, k in enumerate(sorted(metrics_all))}
    return float(loss), {k: float(v) for k, v in metrics_all.items()}


@torch.no_grad()
def eval_model(model_wrapper: ModelTrainerWrapper, tokenizer, val_iter, epoch: int, ignore_index: int, num_candidates: int = 4,
               max_new_tokens: int = 128, seed: Optional[int] = None):
    """reference trainer.py:27-65: caption the first validation image num_candidates times (temperature 0.7, nucleus 0.6).
    Returns the generated ids (prompt stripped); prints decoded text when the tokenizer can decode."""
    model_wrapper.eval()
    device = next(model_wrapper.model.parameters()).device
    images, labels = next(val_iter)
    x = images.to(device)[:1].expand(num_candidates, -1, -1, -1).contiguous()
    label = labels[0]
    prompt = torch.full((num_candidates, 1), tokenizer.bos_token_id, dtype=torch.long, device=device)
    blk = model_wrapper.model.spec["block_size"] - model_wrapper.model.space_for_prompt
    n_new = min(max_new_tokens, blk - prompt.shape[1])
    result = model_wrapper.model.generate(images=x, prompt_ids=prompt, temperature=0.7, max_new_tokens=n_new, nucleus_p=0.6,
                                          seed=seed)[:, 1:]
    _print(f"Model perf at the end of the {epoch}-th epoch")
    if hasattr(tokenizer, "batch_decode"):
        _print("truth", tokenizer.batch_decode([label[label != ignore_index]])[0], "\n")
        for gen in tokenizer.batch_decode(result):
            i = gen.find(tokenizer.eos_token)
            _print(gen[:i] if i >= 0 else gen)
    else:
        _print("truth ids", label[label != ignore_index].tolist())
        for row in result.tolist():
            _print("ids", row[: row.index(tokenizer.eos_token_id)] if tokenizer.eos_token_id in row else row)
    return result


def synthetic_batches(batch_size: int, image_size: int, vocab_size: int, eos: int, seed: int, width: int = 256,
                      pool: int = 8) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """Endless (images, labels) batches of the shapes the reference's loader yields (trainer.py:68-103 +
    WrapperDataLoader): a small pool of pinned synthetic batches cycled forever."""
    batches = []
    for i in range(pool):
        im = synth_images(batch_size, image_size, seed=seed + 2 * i)
        lb = synth_labels(batch_size, width, vocab_size, seed=seed + 2 * i + 1, eos=eos)
        if torch.cuda.is_available():
            im, lb = im.pin_memory(), lb.pin_memory()
        batches.append((im, lb))
    i = 0
    while True:
        yield batches[i % pool]
        i += 1


def main(args) -> dict:
    config = load_training_config(args.config_file)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=device)
    _print(config)
    if config.precision not in ("no", "bf16"):
        raise NotImplementedError(f"precision {config.precision!r}: the B200 path computes in fp32 ('no') or bf16")
    cd = torch.bfloat16 if config.precision == "bf16" else torch.float32
    tokenizer = args.tokenizer
    if tokenizer is None:
        # the hub tokenizer is unreachable offline: GPT-2's special ids (reference trainer.py:116-126 adds missing ones)
        tokenizer = types.SimpleNamespace(eos_token_id=50256, bos_token_id=50256, vocab_size=50257,
                                          mask_token_id=50257 if config.trainer.mask_fraction > 0 else None)
    model_wrapper = ModelTrainerWrapper(config.model, tokenizer, config.trainer, config.ignore_index, device=device,
                                        compute_dtype=cd, spec_overrides=getattr(args, "spec_overrides", None),
                                        seed=getattr(args, "init_seed", None))
    if args.chkpt_file is not None and os.path.exists(args.chkpt_file) and not args.fresh:
        model_wrapper.model.load_partial_checkpoint(args.chkpt_file, map_location=device)     # resume (models/utils.py:31-36)
        model_wrapper.copy_momentum_params()
        _print(f"resumed from {args.chkpt_file}")
    model_wrapper.model.set_dropout_seed(args.seed + rank)
    groups, matchers = build_param_groups(model_wrapper, config)
    for g, oc in zip(groups, config.optimizers):
        _print(f"Optimizing {len(g['params'])} tensors / {sum(p.numel() for p in g['params']):,} parameters with lr={oc.lr} "
               f"and weight_decay={oc.weight_decay}")
    optimizer = (SNRAdam if config.use_snr_optim else AdamW)(groups)
    reducer = GradientAllReducer([p for g in groups for p in g["params"]]) if world > 1 else None
    if reducer is not None:
        reducer.attach_optimizer(optimizer)                 # the 1/world of the gradient mean rides on the fused step's grad_scale
        reducer.broadcast_parameters(model_wrapper.model)
        # the teacher must start from rank 0's student on every rank too (accelerate's DDP wrapper broadcasts the whole
        # ModelTrainerWrapper, model_m included: reference trainer.py:173-174)
        model_wrapper.copy_momentum_params()
    spec = model_wrapper.model.spec
    size = spec["vit_image"]
    vocab = tokenizer.vocab_size
    train_iter = synthetic_batches(config.batch_size, size, vocab, tokenizer.eos_token_id, seed=args.seed + 1000 * rank,
                                   pool=args.pool)
    val_iter = synthetic_batches(config.batch_size, size, vocab, tokenizer.eos_token_id, seed=args.seed + 77 + 1000 * rank, pool=4)
    state, history, train_history = {}, [], []
    for epoch in range(args.epochs if args.epochs is not None else 10000):
        stop = train_loop(model_wrapper, optimizer, train_iter, epoch, config.num_steps if args.steps is None else args.steps,
                          accum=config.gradient_accumulation_steps, reducer=reducer,
                          reset_moco_after_k_epochs=config.reset_moco_after_k_epochs, chckpt_fname=args.chkpt_file,
                          matchers=matchers, graph=bool(args.graph), log_every=args.log_every, state=state)
        if stop:
            break
        train_history.append(state.get("last_train_loss"))
        if args.eval_captions:
            eval_model(model_wrapper, tokenizer, val_iter, epoch, config.ignore_index, max_new_tokens=args.eval_tokens,
                       seed=args.seed)
        loss, metrics = val_loop(model_wrapper, val_iter, epoch, config.num_val_steps if args.val_steps is None else args.val_steps)
        _print(f"Epoch: {epoch}, loss: {loss}, metrics: {metrics}")
        history.append(loss)
    if world > 1:
        torch.distributed.barrier()
    return {"val_losses": history, "train_losses": train_history, "wrapper": model_wrapper}


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config_file", required=True, type=str)
    ap.add_argument("--chkpt_file", required=False, type=str, default=None)
    ap.add_argument("--synthetic", action="store_true", help="synthetic batches (the only built-in data source)")
    ap.add_argument("--epochs", type=int, default=None, help="stop after this many epochs (reference: until the data ends)")
    ap.add_argument("--steps", type=int, default=None, help="override the YAML's num_steps (micro-batches per epoch)")
    ap.add_argument("--val_steps", type=int, default=None, help="override the YAML's num_val_steps")
    ap.add_argument("--graph", type=int, default=0, help="1: micro-step forward + backward as one CUDA-graph replay")
    ap.add_argument("--fresh", action="store_true", help="do not resume from --chkpt_file even if it exists")
    ap.add_argument("--eval_captions", type=int, default=1)
    ap.add_argument("--eval_tokens", type=int, default=32)
    ap.add_argument("--log_every", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--init_seed", type=int, default=None, help="seed of the random initialisation (default: unseeded)")
    ap.add_argument("--pool", type=int, default=8, help="number of distinct synthetic training batches that are cycled")
    args = ap.parse_args(argv)
    args.tokenizer = None
    return args


if __name__ == "__main__":
    main(parse_args())
