"""Fused optimisers behind the reference's optimiser seam (``optim_clazz(param_groups)``, reference trainer.py:169-172).

``AdamW``   -- torch.optim.AdamW semantics (eps 1e-8, amsgrad off, decoupled weight decay) in ONE kernel launch per
               parameter group (libi2t ``i2t_adamw_multi``) instead of a foreach chain per tensor list.
``SNRAdam`` -- the reference's variance-normalised Adam (models/optimizer.py:7-113) in one launch per group
               (``i2t_snradam_multi``) instead of a python loop over parameters.
``ema_update`` -- the momentum-distillation teacher update (training/wrapper.py:53-60), in place, one launch.

State layout matches the originals (``exp_avg`` / ``exp_avg_sq`` / step counter per parameter) so ``state_dict()``
round-trips.  Only parameters whose ``.grad`` is not None are stepped, like the originals (a training loop that keeps
``zero_grad(set_to_none=False)`` for CUDA-graph replays therefore also steps parameters whose gradient is all zero).

The kernels write parameters through raw pointers, which PyTorch cannot see: every step bumps the tensors' version
counters (``torch.autograd.graph.increment_version``) and refreshes, IN PLACE, the bf16 copies the bf16 forward / decode
paths read (``refresh_shadows``) -- their addresses stay valid for captured CUDA graphs.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch.optim.optimizer import Optimizer

from ._lib import call
from .ops import ptr, stream

CHUNK = 1 << 16        # elements per CTA: ~1.8 MB of HBM traffic for AdamW


def shadow_of(p: torch.Tensor):
    """the compute-dtype (bf16) copy of a parameter that VisionEncoderDecoder.weights().c() created, or None"""
    sh = getattr(p, "_i2t_shadow", None)
    if sh is None or sh.device != p.device or sh.shape != p.shape:
        return None
    return sh


class _ShadowRefresher:
    """bf16 copies of a parameter list <- their fp32 masters, one launch (i2t_cast_bf16_multi)."""

    def __init__(self):
        self._table = None

    @torch.no_grad()
    def __call__(self, params: List[torch.Tensor]):
        ps = [p for p in params if shadow_of(p) is not None]
        if not ps:
            return
        if self._table is None:
            self._table = _PointerTable()
        cols = [ps, [shadow_of(p) for p in ps], [None] * len(ps), [None] * len(ps)]
        table, ct, co, cl, n = self._table.get(cols, ps[0].device)
        call("i2t_cast_bf16_multi", ptr(table), ptr(ct), ptr(co), ptr(cl), n, stream())
        for p in ps:
            p._i2t_shadow_version = p._version


class _PointerTable:
    """Device-side [n, 4] pointer table + chunk lists; rebuilt only when a pointer or the tensor set changes."""

    def __init__(self):
        self.sig = None
        self.tabs = None

    def get(self, cols: List[List[torch.Tensor]], device) -> Tuple:
        n = len(cols[0])
        sig = tuple(t.data_ptr() if t is not None else 0 for col in cols for t in col)
        if sig != self.sig:
            rows = [[(cols[c][i].data_ptr() if cols[c][i] is not None else 0) for c in range(4)] for i in range(n)]
            ct, co, cl = [], [], []
            for i in range(n):
                numel = cols[0][i].numel()
                for off in range(0, numel, CHUNK):
                    ct.append(i)
                    co.append(off)
                    cl.append(min(CHUNK, numel - off))
            self.tabs = (torch.tensor(rows, dtype=torch.int64, device=device).contiguous(),
                         torch.tensor(ct, dtype=torch.int32, device=device),
                         torch.tensor(co, dtype=torch.int64, device=device),
                         torch.tensor(cl, dtype=torch.int32, device=device), len(ct))
            self.sig = sig
        return self.tabs


class _FusedAdamBase(Optimizer):
    KERNEL = ""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, eps=1e-8, grad_scale: float = 1.0):
        if lr <= 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if weight_decay < 0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay, eps=eps))
        self.grad_scale = grad_scale
        self._tables: Dict[int, _PointerTable] = {}
        self._refresh = _ShadowRefresher()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stepped: List[torch.Tensor] = []
        for gi, group in enumerate(self.param_groups):
            # parameters are bucketed by their own step count (all equal unless some got no gradient earlier)
            by_step: Dict[int, List[torch.Tensor]] = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("fused optimisers need dense fp32 CUDA parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            b1, b2 = group["betas"]
            for step, ps in by_step.items():
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
                cols = [ps, grads, [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps]]
                tab = self._tables.setdefault((gi, step if len(by_step) > 1 else -1), _PointerTable())
                table, ct, co, cl, n = tab.get(cols, ps[0].device)
                call(self.KERNEL, ptr(table), ptr(ct), ptr(co), ptr(cl), n, float(group["lr"]), float(b1), float(b2),
                     float(group["eps"]), float(group["weight_decay"]), int(step), float(self.grad_scale), stream())
                stepped.extend(ps)
        if stepped:
            torch.autograd.graph.increment_version(stepped)     # the kernels wrote them behind PyTorch's back
            self._refresh(stepped)
        return loss


class AdamW(_FusedAdamBase):
    KERNEL = "i2t_adamw_multi"


class SNRAdam(_FusedAdamBase):
    """Same constructor as the reference's SNRAdam (models/optimizer.py:23-54)."""
    KERNEL = "i2t_snradam_multi"


class EmaUpdater:
    """p_m <- p_m * momentum + p * (1 - momentum) over matching parameter lists, one launch."""

    def __init__(self):
        self._table = _PointerTable()

    @torch.no_grad()
    def __call__(self, params_m: List[torch.Tensor], params: List[torch.Tensor], momentum: float):
        if not params_m:
            return
        # column 2: the teacher weight's bf16 copy (if the bf16 forward created one) is rewritten in the same pass
        cols = [list(params_m), list(params), [shadow_of(p) for p in params_m], [None] * len(params)]
        table, ct, co, cl, n = self._table.get(cols, params_m[0].device)
        call("i2t_ema_multi", ptr(table), ptr(ct), ptr(co), ptr(cl), n, float(momentum), stream())
        torch.autograd.graph.increment_version(list(params_m))
        for p in params_m:
            if getattr(p, "_i2t_shadow", None) is not None:
                p._i2t_shadow_version = p._version
