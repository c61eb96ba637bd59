"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange: after finish(), every rank holds the MEAN over
ranks of the per-rank gradients (the rule of SURVEY 8e / D4), including tied parameters, unused parameters and
gradient accumulation under no_sync()."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(11, 8)
        self.l1 = torch.nn.Linear(8, 16)
        self.l2 = torch.nn.Linear(16, 8)
        self.head = torch.nn.Linear(8, 11, bias=False)
        self.head.weight = self.emb.weight                    # tied like wte / lm_head
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def forward(self, ids):
        h = self.l2(torch.tanh(self.l1(self.emb(ids))))
        return self.head(h).logsumexp(-1).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image2text_b200.dp import GradientAllReducer
    torch.manual_seed(100 + rank)                              # different init per rank: broadcast must fix it
    model = Toy()
    red = GradientAllReducer(model.parameters(), bucket_mb=0.0005)       # tiny buckets -> several all-reduces
    red.broadcast_parameters(model, src=0)
    assert len(red.buckets) >= 3
    g = torch.Generator().manual_seed(7 + rank)
    micro = [torch.randint(0, 11, (4, 5), generator=g) for _ in range(2)]
    # two micro-steps: the first without exchange, the second with
    with red.no_sync():
        model(micro[0]).backward()
    model(micro[1]).backward()
    red.finish()
    got = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reference: single-process autograd on each rank's shard, averaged with plain collectives
    ref = Toy()
    ref.load_state_dict(model.state_dict())
    for m in micro:
        ref(m).backward()
    want = {}
    for n, p in ref.named_parameters():
        if p.grad is not None:
            t = p.grad.clone()
            dist.all_reduce(t)
            want[n] = t / world
    ok = set(got) == set(want) and all(torch.allclose(got[n], want[n], atol=1e-6) for n in want)
    same_init = torch.tensor([float(model.l1.weight.sum())])
    lst = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(lst, same_init)
    ok = ok and bool(lst[0] == lst[1]) and ("unused" not in got)
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
