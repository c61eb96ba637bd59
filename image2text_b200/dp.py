"""Data-parallel gradient exchange: one process per GPU, bucketed all-reduce overlapped with backward.

The reference wraps the trainer in accelerate's DDP (trainer.py:173-174) but then calls ``.module.train_step``
(training/utils.py:76-78), which bypasses ``DDP.forward`` -- DDP's reducer never arms, so no gradient is ever exchanged
and replicas drift.  This module does the exchange the reference intends: the MEAN over ranks of the
per-rank gradients (each rank's loss is already normalised by its local batch, training/wrapper.py:96).

Mechanics: parameters that need a gradient are grouped, in reverse registration order (the order backward produces
them), into flat fp32 buckets of ~``bucket_mb``, and every ``.grad`` IS a view of its bucket (created up front, so autograd
accumulates straight into the buffer NCCL reduces: no pack before and no scatter after the exchange.  The view is bound
the first time a gradient arrives -- a parameter that never receives one keeps ``.grad is None`` and is not stepped, like
in the reference; a ``.grad`` that was replaced behind our back, ``zero_grad(set_to_none=True)``, is re-bound the same way).  ``register_post_accumulate_grad_hook``
counts arrivals; when a bucket is complete an asynchronous ``all_reduce`` is launched on a side stream that waits only on
that point of the compute stream -- the remaining backward kernels keep running.  ``finish()`` joins the side stream; the
1/world of the mean is folded into the fused optimiser's ``grad_scale`` (``attach_optimizer``) or, without one, applied
with one ``mul_`` per bucket.  With
``gradient_accumulation_steps`` > 1 call ``no_sync()`` on the non-final micro-steps.  Collectives: NCCL over
NVLink/NVSwitch on GPUs, gloo on CPU (tests).  Nothing here touches the data path of generation (captions shard by
image, no collective).
"""
from __future__ import annotations

import contextlib
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, process_group=None,
                 only_with_grad: bool = True, graph_events: bool = False):
        # graph_events: CUDA-graphed micro-steps carry one external event-record node per bucket so that the exchange overlaps the
        # backward (see capturing()).  OFF by default: measured on B200 every such node costs ~0.3 ms of graph execution (the nodes
        # cut the programmatic-dependent-launch chain of the replay: gpt2.yaml B=32, 33 buckets: 54.9 -> 66.5 ms of micro-steps at
        # 2 GPUs), while the whole 1.06 GB fp32 exchange takes 2.9 ms over NVLink when finish() runs it after the last replay.
        self.graph_events = graph_events
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        seen, uniq = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params: List[torch.nn.Parameter] = list(reversed(uniq))
        self.buckets: List[List[torch.nn.Parameter]] = []
        cap = int(bucket_mb * 1024 * 1024 / 4)
        cur, cur_n = [], 0
        for p in self.params:
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._flat_views: List[Optional[List[torch.Tensor]]] = [None] * len(self.buckets)
        self.pending = [0] * len(self.buckets)
        self.work = [None] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.sync_enabled = True
        self.side_stream = None
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.only_with_grad = only_with_grad
        self._capturing = False
        self._events: List[Optional["torch.cuda.Event"]] = [None] * len(self.buckets)     # per bucket, recorded INSIDE a graph
        self._scale_in_optimizer = False
        self._reset()
        self._slot = {id(p): (bi, i) for bi, b in enumerate(self.buckets) for i, p in enumerate(b)}

    def attach_optimizer(self, optimizer):
        """Fold the 1/world of the gradient mean into a fused optimiser's `grad_scale` (image2text_b200.optimizer): the kernel
        multiplies the gradient on the fly, so finish() has nothing left to do but wait for the collectives."""
        if self.world > 1 and hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = float(optimizer.grad_scale) / self.world
            self._scale_in_optimizer = True
        return optimizer

    # -- public ------------------------------------------------------------------------------------------------
    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0):
        """DDP-constructor equivalent (reference trainer.py:173-174): rank `src`'s parameters and buffers win."""
        if self.world == 1:
            return
        with torch.no_grad():
            ts = list(module.parameters()) + list(module.buffers())
            for t in ts:
                dist.broadcast(t.data, src=src, group=self.group)
            torch.autograd.graph.increment_version(ts)      # `.data` writes are invisible to the version counters the
            # bf16 weight copies are keyed on (VisionEncoderDecoder.weights)

    @contextlib.contextmanager
    def no_sync(self):
        old = self.sync_enabled
        self.sync_enabled = False
        try:
            yield
        finally:
            self.sync_enabled = old

    # -- CUDA-graphed micro-steps -----------------------------------------------------------------------------------
    # A replayed micro-step runs no Python, so the hooks cannot launch anything during its backward.  Instead, while the step
    # is CAPTURED the hooks record one EXTERNAL event per bucket at the point where the bucket's last gradient has been
    # accumulated (an event-record node of the graph).  After launching the replay of a synchronising micro-step the host
    # queues, per bucket, "wait for that event -> pack -> all-reduce" on the side stream: the exchange of a bucket starts as
    # soon as the running graph passes its record node, i.e. it overlaps the rest of the backward, as in the eager path.
    @contextlib.contextmanager
    def capturing(self):
        """Wrap the capture of a micro-step (forward + backward) so that the hooks record the per-bucket events."""
        self._capturing = True
        self._events = [None] * len(self.buckets)
        self._superseded = []
        self._reset()
        try:
            yield
        finally:
            self._capturing = False
            self._reset()

    def has_graph_events(self) -> bool:
        return any(e is not None for e in self._events)

    def exchange_after_replay(self):
        """Call right after `graph.replay()` of the LAST micro-step of an optimiser step (then `finish()` as usual)."""
        if self.world == 1:
            return
        dev = self.buckets[0][0].device
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=dev)
        for bi, ev in enumerate(self._events):
            if ev is None or self.launched[bi]:
                continue
            self.side_stream.wait_event(ev)
            with torch.cuda.stream(self.side_stream):
                flat = self._pack(bi)
                self.work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.launched[bi] = True

    def finish(self):
        """Call after backward of the last micro-step, before optimizer.step()."""
        if self.world == 1:
            self._reset()
            return
        # buckets whose parameters did not all receive a gradient this step (unused parameters)
        for bi in range(len(self.buckets)):
            if not self.launched[bi] and any(p.grad is not None for p in self.buckets[bi]):
                self._launch(bi)
        for bi, w in enumerate(self.work):
            if w is not None:
                w.wait()
        if self.side_stream is not None:
            torch.cuda.current_stream().wait_stream(self.side_stream)
        inv = 1.0 / self.world
        with torch.no_grad():
            for bi, bucket in enumerate(self.buckets):
                if not self.launched[bi]:
                    continue
                flat, views = self._views(bi)
                if not self._scale_in_optimizer:
                    flat.mul_(inv)                                  # mean over ranks: one launch per bucket
                # gradients that are not views of the bucket (replaced by zero_grad(set_to_none=True) ...): scatter back
                have = [(p.grad, v) for v, p in zip(views, bucket) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
                if have:
                    torch._foreach_copy_([g for g, _ in have], [v for _, v in have])
        self._reset()

    def remove(self):
        for h in self.hooks:
            h.remove()

    # -- internals ---------------------------------------------------------------------------------------------
    def _reset(self):
        for bi, b in enumerate(self.buckets):
            self.pending[bi] = len(b)
            self.work[bi] = None
            self.launched[bi] = False

    def _hook(self, p: torch.nn.Parameter):
        if self.world == 1:
            return
        bi0, i0 = self._slot[id(p)]
        v = self._views(bi0)[1][i0]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():      # first gradient (or a replaced one): move it into the bucket
            with torch.no_grad():
                v.copy_(p.grad)
            p.grad = v
        if self._capturing:
            bi = self.bucket_of[id(p)]
            self.pending[bi] -= 1
            # `<= 0`: a parameter may be announced twice in one backward -- once by a kernel that added its contribution straight
            # into .grad (ops.grad_sinks) and once by AccumulateGrad for a contribution that came through autograd (the tied
            # wte / lm_head weight: LM-head wgrad + embedding gradient).  Every arrival after the count ran out re-records the
            # bucket's event, so the exchange waits for the LAST contribution.
            if self.pending[bi] <= 0:
                ev = torch.cuda.Event(external=True)
                ev.record()                      # an event-record node of the graph being captured
                if self._events[bi] is not None:
                    self._superseded.append(self._events[bi])     # its record node stays in the graph: the event must outlive it
                self._events[bi] = ev
            return
        if not self.sync_enabled:
            return
        bi = self.bucket_of[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _views(self, bi: int):
        """Flat fp32 buffer of bucket `bi` and one view per parameter (created once)."""
        if self.flat[bi] is None:
            bucket = self.buckets[bi]
            n = sum(p.numel() for p in bucket)
            flat = torch.zeros(n, device=bucket[0].device, dtype=torch.float32)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            self.flat[bi] = flat
            self._flat_views[bi] = views
        return self.flat[bi], self._flat_views[bi]

    def _pack(self, bi: int) -> torch.Tensor:
        """Gradients of bucket `bi` -> its flat fp32 buffer (on the current stream): one multi-tensor copy, not one launch
        per parameter (a model that trains every weight has hundreds of them; the launches were most of the exposed time)."""
        flat, views = self._views(bi)
        with torch.no_grad():
            stray = [(v, p.grad) for v, p in zip(views, self.buckets[bi]) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
            gone = [v for v, p in zip(views, self.buckets[bi]) if p.grad is None]
            if gone:
                torch._foreach_zero_(gone)
            if stray:                           # the normal case has nothing to copy: .grad already lives in the bucket
                torch._foreach_copy_([v for v, _ in stray], [g for _, g in stray])
        return flat

    def _launch(self, bi: int):
        bucket = self.buckets[bi]
        dev = bucket[0].device
        flat = self._pack(bi)
        if dev.type == "cuda":
            if self.side_stream is None:
                self.side_stream = torch.cuda.Stream(device=dev)
            self.side_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.side_stream):
                self.work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self.work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched[bi] = True
