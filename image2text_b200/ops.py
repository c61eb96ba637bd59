"""Thin tensor-level wrappers over the C ABI (pointers + sizes in, nothing else).  torch is used for device memory
and streams only; every arithmetic operation below is a libi2t kernel."""
from __future__ import annotations

import contextlib
from typing import NamedTuple, Optional

import torch

from ._lib import call

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2
MASK_NONE, MASK_CAUSAL, MASK_PROMPT = 0, 1, 2


class DropSite(NamedTuple):
    """One dropout call of a training forward pass: probability, the device rng state (uint64 {seed, step offset},
    stored as an int64[2] tensor) and the site index.  The mask is a pure function of these (csrc/rng.cuh)."""
    p: float
    state: torch.Tensor
    site: int


class DropCtx:
    """Hands out consecutive site indices over one forward pass.  `base` separates passes that run on the same step
    offset (student forward = 0, EMA-teacher forward = 1 << 20)."""

    def __init__(self, state: torch.Tensor, base: int = 0):
        assert state.dtype == torch.int64 and state.numel() == 2 and state.is_cuda
        self.state = state
        self.base = base
        self.n = 0

    def site(self, p: float) -> Optional[DropSite]:
        """Next site; None when p == 0 (the index is consumed either way so site numbering does not depend on p)."""
        i = self.base + self.n
        self.n += 1
        return DropSite(float(p), self.state, i) if p > 0.0 else None


class _GradSinks:
    """State of `grad_sinks()`; `uses` counts the backward nodes that still owe a contribution to a parameter."""
    on = False
    notify = None
    uses: dict = {}


@contextlib.contextmanager
def grad_sinks(notify=None):
    """Inside this context the backward kernels of the parameter-consuming ops (linear / Conv1D weight and bias gradients,
    LayerNorm gamma / beta, position embeddings, the LSH EmbeddingBags) ADD their result straight into the parameter's
    existing fp32 ``.grad`` buffer and hand autograd no gradient for it: no temporary, no zero fill, no AccumulateGrad ``add_``
    and no ``slice_backward`` for row views of a packed weight (nn.MultiheadAttention.in_proj_weight).  The arithmetic is the one
    autograd would do (grad += contribution).  AccumulateGrad hooks do not fire for such parameters; `notify(param)` is called
    instead, once per parameter, when its last contribution has been added (the data-parallel reducer's hook).  Used while a
    micro-step is captured into a CUDA graph (wrapper.train_step_graphed), where ``.grad`` buffers already exist and are static."""
    st = _GradSinks
    prev = (st.on, st.notify, st.uses)
    st.on, st.notify, st.uses = True, notify, {}
    try:
        yield
    finally:
        st.on, st.notify, st.uses = prev


def sink_use(p, count: bool = True) -> bool:
    """Forward-time: will the gradient contribution of this use of parameter `p` be added into ``p.grad`` in place?"""
    st = _GradSinks
    if not st.on or p is None or not isinstance(p, torch.nn.Parameter) or not p.requires_grad:
        return False
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return False
    if count:
        st.uses[id(p)] = st.uses.get(id(p), 0) + 1
    return True


def sink_done(p):
    """Backward-time: one contribution to ``p.grad`` has been queued; after the last one the reducer is told."""
    st = _GradSinks
    left = st.uses.get(id(p), 1) - 1
    st.uses[id(p)] = left
    if left == 0 and st.notify is not None:
        st.notify(p)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def tdt(code_or_dtype) -> torch.dtype:
    return {F32: torch.float32, BF16: torch.bfloat16}.get(code_or_dtype, code_or_dtype)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("image2text_b200 kernels need CUDA tensors (there is no CPU fallback)")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: Optional[torch.Tensor], eps: float, out_dtype=torch.float32,
              rows: Optional[int] = None, row_stride: Optional[int] = None, want_stats: bool = False):
    """x: (..., C) contiguous unless rows/row_stride are given (strided row view of a larger buffer)."""
    _need_cuda(x, gamma)
    C = x.shape[-1]
    if rows is None:
        assert x.is_contiguous()
        rows = x.numel() // C
        row_stride = C
        out_shape = x.shape
    else:
        out_shape = (rows, C)
    y = torch.empty(out_shape, device=x.device, dtype=out_dtype)
    mean = rstd = None
    if want_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    call("i2t_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, C, row_stride, eps, dt(x),
         dt(y), stream())
    return (y, mean, rstd) if want_stats else y


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dx_dtype=torch.float32, dx_add=None):
    """dx (+ dx_add: the gradient that reaches x over the skip connection, same dtype / shape as dx)."""
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty(x.shape, device=x.device, dtype=dx_dtype)
    if dx_add is None:
        call("i2t_layernorm_bwd", ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(dgamma), ptr(dbeta), rows, C,
             dt(dy), dt(x), dt(dx), stream())
    else:
        assert dx_add.dtype == dx_dtype and dx_add.is_contiguous() and dx_add.numel() == dx.numel()
        call("i2t_layernorm_bwd_add", ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx_add), ptr(dx), ptr(dgamma),
             ptr(dbeta), rows, C, dt(dy), dt(x), dt(dx), stream())
    return dx


def gemm(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, out_dtype=torch.float32, a_kmajor: bool = True,
         b_kmajor: bool = True, accumulate: bool = False, M=None, N=None, K=None, lda=None, ldb=None, ldc=None,
         b_stable: bool = False):
    """C[M,N] = act(op(A) op(B) + bias) + residual.  Default: A (M,K) row-major, B (N,K) row-major (nn.Linear weight).
    a/b must be 2-D with a contiguous inner dimension (a row pitch is allowed)."""
    _need_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and a.dtype == b.dtype
    if M is None:
        M, K = (a.shape if a_kmajor else (a.shape[1], a.shape[0]))
        N = b.shape[0] if b_kmajor else b.shape[1]
        Kb = b.shape[1] if b_kmajor else b.shape[0]
        assert K == Kb, (a.shape, b.shape)
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.stride(-1) == 1
    ldc = (out.stride(0) if out.dim() == 2 else N) if ldc is None else ldc
    if residual is not None:
        assert residual.stride(-1) == 1 and (residual.stride(0) if residual.dim() == 2 else N) == ldc
    call("i2t_gemm_ex", ptr(a), ptr(b), ptr(bias), ptr(residual), ptr(out), M, N, K, lda, ldb, ldc, int(a_kmajor),
         int(b_kmajor), act, int(accumulate), dt(a), dt(residual) if residual is not None else F32, dt(out), int(b_stable),
         stream())
    return out


def colsum_(x2d: torch.Tensor, out: torch.Tensor):
    call("i2t_colsum", ptr(x2d), ptr(out), x2d.shape[0], x2d.shape[1], x2d.stride(0), dt(x2d), stream())


def _attn_fwd(drop: Optional[DropSite], *args):
    if drop is None:
        call("i2t_attn_fwd", *args, stream())
    else:
        call("i2t_attn_fwd_dropout", *args, drop.p, ptr(drop.state), drop.site, stream())


def _attn_bwd(drop: Optional[DropSite], *args):
    if drop is None:
        call("i2t_attn_bwd", *args, stream())
    else:
        call("i2t_attn_bwd_dropout", *args, drop.p, ptr(drop.state), drop.site, stream())


def attention_packed(qkv: torch.Tensor, B: int, T: int, H: int, mask_mode: int, n_prompt: int = 0, want_lse: bool = False,
                     drop: Optional[DropSite] = None):
    """qkv: (B*T, 3C) packed [q | k | v] as c_attn / in_proj produce it.  Returns (B*T, C) in qkv's dtype."""
    C = qkv.shape[1] // 3
    hs = C // H
    out = torch.empty((B * T, C), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32) if want_lse else None
    es = qkv.element_size()
    base = qkv.data_ptr()
    _attn_fwd(drop, base, base + C * es, base + 2 * C * es, ptr(out), ptr(lse), B, H, T, T, hs, T * 3 * C, 3 * C,
              T * 3 * C, 3 * C, mask_mode, n_prompt, dt(qkv), dt(out))
    return (out, lse) if want_lse else out


def attention_packed_bwd(qkv, out, dout, lse, B, T, H, mask_mode, n_prompt, drop: Optional[DropSite] = None,
                         tok: Optional[DropSite] = None):
    """`tok`: the token-level q / k / v dropout site the forward applied to `qkv` (token_dropout_): the returned gradient of the packed
    buffer is already multiplied by those masks (by the tcgen05 kernel on its way out, or by a trailing pass)."""
    C = qkv.shape[1] // 3
    hs = C // H
    dqkv = torch.empty_like(qkv)
    from ._lib import lib
    ws_bytes = lib().i2t_attn_bwd_workspace_bytes(B, H, T, hs)
    ws = torch.empty(ws_bytes, device=qkv.device, dtype=torch.uint8)
    es = qkv.element_size()
    b0, d0 = qkv.data_ptr(), dqkv.data_ptr()
    if tok is not None:
        assert drop is None or drop.state is tok.state
        call("i2t_attn_bwd_dropout_tok", b0, b0 + C * es, b0 + 2 * C * es, ptr(out), ptr(dout), ptr(lse), d0, d0 + C * es, d0 + 2 * C * es,
             ptr(ws), B, H, T, hs, T * 3 * C, 3 * C, mask_mode, n_prompt, dt(qkv), drop.p if drop is not None else 0.0, ptr(tok.state),
             drop.site if drop is not None else 0, tok.p, tok.site, stream())
        return dqkv
    _attn_bwd(drop, b0, b0 + C * es, b0 + 2 * C * es, ptr(out), ptr(dout), ptr(lse), d0, d0 + C * es, d0 + 2 * C * es,
              ptr(ws), B, H, T, T, hs, T * 3 * C, 3 * C, T * 3 * C, 3 * C, mask_mode, n_prompt, dt(qkv))
    return dqkv


def xattn(q: torch.Tensor, kv: torch.Tensor, B: int, T: int, S: int, H: int):
    """q: (B*T, C); kv: (B*S, 2C) packed [k | v].  Returns (B*T, C)."""
    C = q.shape[1]
    hs = C // H
    out = torch.empty_like(q)
    es = kv.element_size()
    call("i2t_xattn_fwd", ptr(q), kv.data_ptr(), kv.data_ptr() + C * es, ptr(out), B, H, T, S, hs, q.stride(0), S * 2 * C,
         2 * C, dt(q), dt(out), stream())
    return out


def xattn_bwd(q, kv, dout, B, T, S, H):
    C = q.shape[1]
    hs = C // H
    dq = torch.empty_like(q)
    dkv = torch.zeros((B * S, 2 * C), device=q.device, dtype=torch.float32)
    es = kv.element_size()
    call("i2t_xattn_bwd", ptr(q), kv.data_ptr(), kv.data_ptr() + C * es, ptr(dout), ptr(dq), dkv.data_ptr(),
         dkv.data_ptr() + C * 4, B, H, T, S, hs, q.stride(0), S * 2 * C, 2 * C, S * 2 * C, 2 * C, dt(q), stream())
    return dq, dkv


def xattn_tc(q: torch.Tensor, kv: torch.Tensor, B: int, T: int, S: int, H: int, drop: Optional[DropSite] = None):
    """Cross attention on the general attention kernels (i2t_attn_fwd with Tk = S, no mask; tensor cores for bf16):
    q (B*T, C), kv (B*S, 2C) packed [k | v].  Returns (out (B*T, C), lse (B, H, T))."""
    C = q.shape[1]
    hs = C // H
    out = torch.empty_like(q)
    lse = torch.empty((B, H, T), device=q.device, dtype=torch.float32)
    es = kv.element_size()
    _attn_fwd(drop, ptr(q), kv.data_ptr(), kv.data_ptr() + C * es, ptr(out), ptr(lse), B, H, T, S, hs, T * q.stride(0),
              q.stride(0), S * 2 * C, 2 * C, MASK_NONE, 0, dt(q), dt(out))
    return out, lse


def xattn_tc_bwd(q, kv, out, dout, lse, B, T, S, H, drop: Optional[DropSite] = None):
    C = q.shape[1]
    hs = C // H
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    from ._lib import lib
    ws = torch.empty(lib().i2t_attn_bwd_workspace_bytes(B, H, T, hs), device=q.device, dtype=torch.uint8)
    es = kv.element_size()
    _attn_bwd(drop, ptr(q), kv.data_ptr(), kv.data_ptr() + C * es, ptr(out), ptr(dout), ptr(lse), ptr(dq), dkv.data_ptr(),
              dkv.data_ptr() + C * es, ptr(ws), B, H, T, S, hs, T * q.stride(0), q.stride(0), S * 2 * C, 2 * C, MASK_NONE, 0,
              dt(q))
    return dq, dkv


def dropout_add(y: torch.Tensor, residual: Optional[torch.Tensor], drop: DropSite) -> torch.Tensor:
    """residual + dropout(y) as one pass (fp32 out); y in the compute dtype."""
    assert y.is_contiguous() and (residual is None or (residual.is_contiguous() and residual.dtype == torch.float32))
    out = torch.empty(y.shape, device=y.device, dtype=torch.float32)
    call("i2t_dropout_add_fwd", ptr(y), ptr(residual), ptr(out), y.numel(), drop.p, ptr(drop.state), drop.site, dt(y), stream())
    return out


def dropout_bwd(dy: torch.Tensor, out_dtype, drop: DropSite) -> torch.Tensor:
    """keep * dy / (1-p), written in `out_dtype` (the cast the linear's backward needs anyway)."""
    assert dy.is_contiguous()
    g = torch.empty(dy.shape, device=dy.device, dtype=out_dtype)
    call("i2t_dropout_bwd", ptr(dy), ptr(g), dy.numel(), drop.p, ptr(drop.state), drop.site, dt(dy), dt(g), stream())
    return g


def token_dropout_(x2d: torch.Tensor, seg: int, nseg: int, drop: DropSite) -> torch.Tensor:
    """In place: one Bernoulli per (row, segment) of a packed (rows, nseg*seg) buffer (reference models/layers.py:454-461)."""
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    call("i2t_token_dropout", ptr(x2d), x2d.shape[0], x2d.stride(0), seg, nseg, drop.p, ptr(drop.state), drop.site, dt(x2d),
         stream())
    return x2d


def rng_advance(state: torch.Tensor):
    call("i2t_rng_advance", ptr(state), stream())


def patch_im2col(images: torch.Tensor, p: int, out_dtype) -> torch.Tensor:
    _need_cuda(images)
    assert images.dtype == torch.float32 and images.is_contiguous()
    B, _, H, W = images.shape
    out = torch.empty((B * (H // p) * (W // p), 3 * p * p), device=images.device, dtype=out_dtype)
    call("i2t_patch_im2col", ptr(images), ptr(out), B, H, W, p, dt(out), stream())
    return out


def vit_assemble(patch_out: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, B: int) -> torch.Tensor:
    npatch = patch_out.shape[0] // B
    C = patch_out.shape[1]
    x = torch.empty((B, npatch + 1, C), device=patch_out.device, dtype=torch.float32)
    call("i2t_vit_assemble", ptr(patch_out), ptr(cls), ptr(pos), ptr(x), B, npatch, C, dt(patch_out), stream())
    return x


def embed(ids: Optional[torch.Tensor], prompt: Optional[torch.Tensor], wte, wpe, B: int, T: int, n_prompt: int, S: int):
    C = wte.shape[1]
    x = torch.empty((B, T, C), device=wte.device, dtype=torch.float32)
    call("i2t_embed_fwd", ptr(ids), ptr(prompt), ptr(wte), ptr(wpe), ptr(x), B, T, n_prompt, S, C, stream())
    return x
