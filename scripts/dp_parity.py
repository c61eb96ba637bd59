"""N-GPU data-parallel parity: after the overlapped all-reduce every rank's gradients equal the MEAN over ranks
of the single-process gradients of each rank's shard (tiny config, fp32, dropout 0).  Run under torchrun."""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import load_training_config  # noqa: E402
from image2text_b200.config_schema import TrainerWrapperConfig  # noqa: E402
from image2text_b200.dp import GradientAllReducer  # noqa: E402
from image2text_b200.model_spec import synth_state_dict  # noqa: E402
from image2text_b200.synthetic import synth_images, synth_labels  # noqa: E402
from image2text_b200.wrapper import ModelTrainerWrapper  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tc = load_training_config(os.path.join(ROOT, "configs", "tiny.yaml"))
tok = types.SimpleNamespace(eos_token_id=612, bos_token_id=612, mask_token_id=None, vocab_size=613)
over = dict(vit_layers=2, vit_image=32)


def grads_for(shard_rank, reducer_on):
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(), -100, device=f"cuda:{local}", spec_overrides=over)
    w.model.load_state_dict(synth_state_dict(w.model.spec, seed=0))
    w.train()
    red = GradientAllReducer(w.model.parameters(), bucket_mb=0.25) if reducer_on else None
    images = synth_images(3, 32, seed=100 + shard_rank).cuda()
    labels = synth_labels(3, 20, 613, seed=200 + shard_rank, min_len=3, max_len=14, eos=612).cuda()
    loss, _ = w.train_step(images, labels)
    loss.backward()
    if red is not None:
        red.finish()
        nb = len(red.buckets)
        red.remove()
    else:
        nb = 0
    return {n: p.grad.clone() for n, p in w.model.named_parameters() if p.grad is not None}, nb


got, nb = grads_for(rank, True)
want = None
for r in range(world):                       # every rank recomputes every shard single-process: no collective involved
    g, _ = grads_for(r, False)
    want = g if want is None else {k: want[k] + g[k] for k in g}
worst = 0.0
for k in want:
    ref = want[k] / world
    err = float((got[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
    worst = max(worst, err)
ok = worst < 1e-5 and set(got) == set(want)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DP parity world={world}: buckets={nb} tensors={len(want)} worst rel err={worst:.2e} -> {'OK' if flag.item() == 1 else 'FAIL'}")


def grads_graphed(shard_rank, micro=5, graph_events=True):
    """CUDA-graphed micro-steps (calls 1-2 eager, 3 captures, 3-5 replay): the exchange of the last replay is queued behind
    the per-bucket events recorded inside the graph (GradientAllReducer.exchange_after_replay)."""
    w = ModelTrainerWrapper(tc.model, tok, TrainerWrapperConfig(), -100, device=f"cuda:{local}", spec_overrides=over)
    w.model.load_state_dict(synth_state_dict(w.model.spec, seed=0))
    w.train()
    red = GradientAllReducer(w.model.parameters(), bucket_mb=0.25, graph_events=graph_events)
    images = synth_images(3, 32, seed=100 + shard_rank).cuda()
    labels = synth_labels(3, 20, 613, seed=200 + shard_rank, min_len=3, max_len=14, eos=612).cuda()
    for i in range(micro):
        with red.no_sync():
            w.train_step_graphed(images, labels, 1.0, reducer=red, sync=i == micro - 1)
    n_ev = sum(e is not None for e in red._events)
    red.finish()
    red.remove()
    return {n: p.grad.clone() for n, p in w.model.named_parameters() if p.grad is not None}, n_ev


got_g, n_ev = grads_graphed(rank)
worst_g = 0.0
for k in want:
    ref = 5.0 * want[k] / world
    worst_g = max(worst_g, float((got_g[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-20)))
ok_g = worst_g < 1e-5 and set(got_g) == set(want) and n_ev > 0
flag = torch.tensor([1.0 if ok_g else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DP parity (graphed micro-steps, exchange behind in-graph events) world={world}: bucket events={n_ev} "
          f"worst rel err={worst_g:.2e} -> {'OK' if flag.item() == 1 else 'FAIL'}")
all_ok = flag.item() == 1
# default reducer: no in-graph events, every bucket exchanged by finish() after the last replay
got_e, n_ev_e = grads_graphed(rank, graph_events=False)
worst_e = 0.0
for k in want:
    ref = 5.0 * want[k] / world
    worst_e = max(worst_e, float((got_e[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-20)))
ok_e = worst_e < 1e-5 and set(got_e) == set(want) and n_ev_e == 0
flag = torch.tensor([1.0 if ok_e else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DP parity (graphed micro-steps, exchange after the last replay) world={world}: worst rel err={worst_e:.2e} "
          f"-> {'OK' if flag.item() == 1 else 'FAIL'}")
all_ok = all_ok and flag.item() == 1
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)
