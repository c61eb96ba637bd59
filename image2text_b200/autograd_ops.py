"""torch.autograd.Function wrappers: the forward AND backward of every op is a libi2t kernel; autograd only records the
graph (it is plumbing -- PyTorch computes nothing here).  Activations flow in the compute dtype (fp32, or bf16 with the
autocast semantics of SURVEY.md Q9: bf16 linears/attention, fp32 LayerNorm / residual stream / loss / master grads)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import call
from .ops import dt, ptr, stream


def _ln_forward(ctx, x, gamma, beta, eps, out_dtype):
    x = x.contiguous()
    need = any(ctx.needs_input_grad)
    if need:
        y, mean, rstd = ops.layernorm(x, gamma, beta, eps, out_dtype, want_stats=True)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.has_beta = beta is not None
        # ops.grad_sinks(): the kernel accumulates (+=) d gamma / d beta, so it can write the .grad buffers themselves
        ctx.sinks = None
        if ctx.needs_input_grad[1] and ops.sink_use(gamma, False) and (beta is None or ops.sink_use(beta, False)):
            ctx.sinks = (gamma, beta)
            ops.sink_use(gamma), ops.sink_use(beta)
    else:
        y = ops.layernorm(x, gamma, beta, eps, out_dtype)
    return x, y


def _ln_backward(ctx, dy, dskip=None):
    x, gamma, mean, rstd = ctx.saved_tensors
    dy = dy.contiguous()
    if dskip is not None:
        dskip = dskip.contiguous()
        if dskip.dtype != x.dtype:
            dskip = dskip.to(x.dtype)
    want_g = ctx.needs_input_grad[1]
    if ctx.sinks is not None:
        pg, pb = ctx.sinks
        dx = ops.layernorm_bwd(dy, x, gamma, mean, rstd, pg.grad, pb.grad if pb is not None else None, dx_dtype=x.dtype,
                               dx_add=dskip)
        ops.sink_done(pg)
        if pb is not None:
            ops.sink_done(pb)
        return dx, None, None, None, None
    dgamma = torch.zeros_like(gamma, dtype=torch.float32) if want_g else None
    dbeta = torch.zeros_like(gamma, dtype=torch.float32) if (want_g and ctx.has_beta) else None
    dx = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dx_dtype=x.dtype, dx_add=dskip)
    return dx, dgamma, dbeta, None, None


class LayerNormFn(torch.autograd.Function):
    """F.layer_norm over the last dim: reference models/layers.py:357-358 (eps 1e-5), torchvision eps 1e-6."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype):
        return _ln_forward(ctx, x, gamma, beta, eps, out_dtype)[1]

    @staticmethod
    def backward(ctx, dy):
        return _ln_backward(ctx, dy)


class LayerNormSkipFn(torch.autograd.Function):
    """The LayerNorm of a pre-LN block together with the skip connection that leaves the same tensor (x feeds ln_k AND the
    residual add, models/layers.py:597-606): returns (LayerNorm(x), x).  The gradient of x is the sum over both paths; taking
    both through one node lets the backward kernel add the skip gradient in its single pass over the rows instead of autograd
    running one more elementwise add per residual join."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype):
        x, y = _ln_forward(ctx, x, gamma, beta, eps, out_dtype)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        if dy is None:                     # the normalised branch went unused
            return dskip, None, None, None, None
        return _ln_backward(ctx, dy, dskip)


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) + residual with W an nn.Linear weight (N,K): reference models/layers.py:452,469,482,484,
    models/decoder.py:256, nn.MultiheadAttention projections.  `w_c` is the compute-dtype view of the master weight.
    Training-mode dropout (ops.DropSite): `drop` = nn.Dropout on the projection before the residual add
    (models/layers.py:469,485): y = residual + dropout(x W^T + b); `tok_drop` = the per-token q/k/v masks of
    models/layers.py:454-461 applied to the packed c_attn output."""

    @staticmethod
    def forward(ctx, x2d, w, w_c, bias, residual, act, out_dtype, pad_rows=False, drop=None, tok_drop=None, sink=None):
        need = any(ctx.needs_input_grad)      # grad mode is always off inside Function.forward
        out = None
        # an fp32 input of a bf16 linear (the final hidden state in front of the LM head) is cast HERE, not by an autograd node of
        # its own: the data gradient then leaves the dgrad GEMM in fp32 directly (split-K eligible: K = 50257 at M = B*T rows),
        # with no bf16 round trip and no cast kernel in the backward
        x_dtype = x2d.dtype
        if x2d.dtype != w_c.dtype:
            x2d = x2d.to(w_c.dtype)
        if pad_rows and w_c.shape[0] % 8 != 0:
            # rows padded to a multiple of 8 elements (e.g. V = 50257 -> pitch 50264): the output is a strided view, and
            # the gradient that comes back with the same pitch satisfies TMA's 16-byte row-pitch rule for dgrad / wgrad
            N = w_c.shape[0]
            out = torch.empty((x2d.shape[0], (N + 7) // 8 * 8), device=x2d.device, dtype=out_dtype)[:, :N]
        # B operand = a parameter (or its bf16 shadow): nothing between two optimiser steps writes it, and the kernels that do
        # (optimiser, EMA, shadow refresh) never trigger their dependents early -- the GEMM may fetch its first weight tiles
        # before the previous kernel has finished (I2T_GEMM_B_STABLE)
        bs = isinstance(w, torch.nn.Parameter)
        if drop is not None:
            assert act == ops.ACT_NONE and out is None and out_dtype == torch.float32
            z = None
            y = ops.dropout_add(ops.gemm(x2d, w_c, bias=bias, out_dtype=x2d.dtype, b_stable=bs), residual, drop)
        elif need and act != ops.ACT_NONE:
            z = ops.gemm(x2d, w_c, bias=bias, out_dtype=x2d.dtype, b_stable=bs)
            y = torch.empty(z.shape, device=z.device, dtype=out_dtype)
            call("i2t_act_fwd", ptr(z), ptr(y), z.numel(), act, dt(z), dt(y), stream())
            assert residual is None
        else:
            z = None
            y = ops.gemm(x2d, w_c, bias=bias, residual=residual, act=act, out_dtype=out_dtype, out=out, b_stable=bs)
        if tok_drop is not None:
            site, seg, nseg = tok_drop[:3]
            ops.token_dropout_(y, seg, nseg, site)
        if need:
            ctx.save_for_backward(x2d, w_c, z)
            ctx.act = act
            ctx.has_bias = bias is not None
            ctx.has_res = residual is not None
            ctx.res_dtype = residual.dtype if residual is not None else None
            ctx.drop, ctx.tok_drop = drop, tok_drop
            ctx.b_stable = bs
            ctx.x_dtype = x_dtype
            ctx.sink_w, ctx.sink_b = _sinks_of(ctx, w, bias, sink)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, w_c, z = ctx.saved_tensors
        if not (dy.dim() == 2 and dy.stride(1) == 1 and dy.stride(0) % 8 == 0 and dy.stride(0) >= dy.shape[1]):
            dy = dy.contiguous()      # row-padded gradients (LM head) are consumed in place
        dres = dy.to(ctx.res_dtype) if (ctx.has_res and ctx.needs_input_grad[4]) else None
        g = _grad_in(dy, x2d.dtype, ctx.drop, ctx.tok_drop)
        if z is not None:
            dz = torch.empty_like(z)
            call("i2t_act_bwd", ptr(z), ptr(g), ptr(dz), z.numel(), ctx.act, dt(z), dt(g), stream())
            g = dz
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dX[M,K] = dY[M,N] W[N,K]
            dx = ops.gemm(g, w_c, out_dtype=ctx.x_dtype, a_kmajor=True, b_kmajor=False, b_stable=ctx.b_stable)
        if ctx.needs_input_grad[1]:
            # dW[N,K] = dY^T[N,M] X[M,K]  (fp32 master gradient)
            M, N = g.shape
            K = x2d.shape[1]
            if ctx.sink_w is not None:             # ops.grad_sinks(): W.grad[rows] += dY^T X, nothing for autograd to add
                pw, rows = ctx.sink_w
                ops.gemm(g, x2d, out=pw.grad if rows is None else pw.grad[rows], accumulate=True, a_kmajor=False, b_kmajor=False,
                         M=N, N=K, K=M, lda=g.stride(0), ldb=x2d.stride(0), ldc=K)
                ops.sink_done(pw)
            else:
                dw = torch.empty((N, K), device=g.device, dtype=torch.float32)
                ops.gemm(g, x2d, out=dw, a_kmajor=False, b_kmajor=False, M=N, N=K, K=M, lda=g.stride(0), ldb=x2d.stride(0),
                         ldc=K)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = _bias_grad(ctx, g)
        return dx, dw, None, db, dres, None, None, None, None, None, None


def _sinks_of(ctx, w, bias, sink):
    """(parameter, rows) targets of the weight / bias gradient under ops.grad_sinks(), else (None, None).  `sink` =
    (weight parameter, row slice, bias parameter) when w / bias are row views of packed parameters."""
    pw, rows, pb = sink if sink is not None else (w, None, bias)
    sw = (pw, rows) if (ctx.needs_input_grad[1] and ops.sink_use(pw)) else None
    sb = (pb, rows) if (bias is not None and ctx.needs_input_grad[3] and ops.sink_use(pb)) else None
    return sw, sb


def _bias_grad(ctx, g):
    """d bias = column sums of the output gradient (i2t_colsum accumulates: into .grad under ops.grad_sinks())."""
    if ctx.sink_b is not None:
        pb, rows = ctx.sink_b
        ops.colsum_(g, pb.grad if rows is None else pb.grad[rows])
        ops.sink_done(pb)
        return None
    db = torch.zeros(g.shape[1], device=g.device, dtype=torch.float32)
    ops.colsum_(g, db)
    return db


def _grad_in(dy, cd, drop, tok_drop):
    """The gradient w.r.t. the GEMM output in the compute dtype: the dropout multiplier and the autocast cast are one
    pass; the token-level q/k/v masks scale the (freshly produced) gradient of the packed buffer in place."""
    if drop is not None:
        g = ops.dropout_bwd(dy.contiguous(), cd, drop)
    else:
        g = dy if dy.dtype == cd else dy.to(cd)      # autocast: the linear's grad flows in the activation dtype
    if tok_drop is not None and not (len(tok_drop) > 3 and tok_drop[3]):
        site, seg, nseg = tok_drop[:3]  # dy is the attention backward's freshly written dqkv: scaled in place
        ops.token_dropout_(g, seg, nseg, site)       # (4th element set: the attention backward has applied the masks already)
    return g


def linear(x2d, w, w_c, bias=None, residual=None, act=ops.ACT_NONE, out_dtype=torch.float32, pad_rows=False, drop=None,
           tok_drop=None, sink=None):
    return LinearFn.apply(x2d, w, w_c, bias, residual, act, out_dtype, pad_rows, drop, tok_drop, sink)


class AttnFn(torch.autograd.Function):
    """Self attention on a packed (B*T, 3C) qkv buffer: reference models/layers.py:465 (+ the mask algebra of
    vision_encoder_decoder.py:75-111 / layers.py:581-595 folded into `mask_mode`), torchvision :113.
    `drop`: dropout_p of the SDPA call (training mode), regenerated -- not stored -- by the backward kernel.
    `tok`: the token-level q / k / v dropout site already applied to `qkv` by the producing linear; the backward then returns the
    gradient of the packed buffer multiplied by those masks (the linear skips its own masking pass)."""

    @staticmethod
    def forward(ctx, qkv, B, T, H, mask_mode, n_prompt, drop=None, tok=None):
        qkv = qkv.contiguous()
        if any(ctx.needs_input_grad):
            out, lse = ops.attention_packed(qkv, B, T, H, mask_mode, n_prompt, want_lse=True, drop=drop)
            ctx.save_for_backward(qkv, out, lse)
            ctx.meta = (B, T, H, mask_mode, n_prompt, drop)
            ctx.tok = tok
        else:
            out = ops.attention_packed(qkv, B, T, H, mask_mode, n_prompt, drop=drop)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        B, T, H, mask_mode, n_prompt, drop = ctx.meta
        dqkv = ops.attention_packed_bwd(qkv, out, dout.contiguous().to(qkv.dtype), lse, B, T, H, mask_mode, n_prompt, drop=drop,
                                        tok=ctx.tok)
        return dqkv, None, None, None, None, None, None, None


class XAttnFn(torch.autograd.Function):
    """Cross attention over S <= 64 encoder tokens: reference models/layers.py:537-542,600-605.  bf16: the tensor-core
    attention kernels with Tk = S (one key tile); fp32: the shared-memory K/V kernels of xattn.cu (parity anchor).
    With dropout on the probabilities (nn.MultiheadAttention(dropout=), HF attn_pdrop) both dtypes run the general
    attention kernels, which regenerate the mask in the backward."""

    @staticmethod
    def forward(ctx, q, kv, B, T, S, H, drop=None):
        q, kv = q.contiguous(), kv.contiguous()
        tc = (q.dtype == torch.bfloat16 or drop is not None) and (q.shape[1] // H) in (32, 64) and q.shape[1] % 8 == 0
        if drop is not None and not tc:
            raise RuntimeError("cross-attention dropout needs head_dim 32 or 64")
        if tc:
            out, lse = ops.xattn_tc(q, kv, B, T, S, H, drop=drop)
        else:
            out, lse = ops.xattn(q, kv, B, T, S, H), None
        if any(ctx.needs_input_grad):
            if tc:
                ctx.save_for_backward(q, kv, out, lse)
            else:
                ctx.save_for_backward(q, kv)
            ctx.meta = (B, T, S, H, tc, drop)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, S, H, tc, drop = ctx.meta
        if tc:
            q, kv, out, lse = ctx.saved_tensors
            dq, dkv = ops.xattn_tc_bwd(q, kv, out, dout.contiguous().to(q.dtype), lse, B, T, S, H, drop=drop)
            return dq, dkv, None, None, None, None, None
        q, kv = ctx.saved_tensors
        dq, dkv = ops.xattn_bwd(q, kv, dout.contiguous().to(q.dtype), B, T, S, H)
        return dq, dkv.to(kv.dtype), None, None, None, None, None


class EmbedFn(torch.autograd.Function):
    """cat(prompt, wte[ids])[:T] + wpe[:T]: reference vision_encoder_decoder.py:84-88 + decoder.py:234-243."""

    @staticmethod
    def forward(ctx, ids, prompt, wte, wpe, T, n_prompt, drop=None):
        B, S = ids.shape
        x = ops.embed(ids, prompt, wte, wpe, B, T, n_prompt, S)
        if drop is not None:                                  # transformer.drop, models/decoder.py:236-243
            x = ops.dropout_add(x, None, drop)
        ctx.save_for_backward(ids)
        ctx.meta = (B, T, n_prompt, S, wte.shape, prompt is not None, wpe.shape[0], drop)
        ctx.sink_wpe = wpe if (ctx.needs_input_grad[3] and ops.sink_use(wpe)) else None
        return x

    @staticmethod
    def backward(ctx, dx):
        (ids,) = ctx.saved_tensors
        B, T, n_prompt, S, wte_shape, has_prompt, wpe_rows, drop = ctx.meta
        dx = dx.contiguous()
        if drop is not None:
            dx = ops.dropout_bwd(dx, torch.float32, drop)
        C = dx.shape[-1]
        dprompt = dwte = dwpe = None
        if has_prompt and ctx.needs_input_grad[1]:
            dprompt = dx[:, :n_prompt].contiguous()
        if ctx.needs_input_grad[2]:
            dwte = torch.zeros(wte_shape, device=dx.device, dtype=torch.float32)
            call("i2t_embed_bwd", ptr(ids), ptr(dx), ptr(dwte), B, T, n_prompt, S, C, stream())
        if ctx.sink_wpe is not None:               # ops.grad_sinks(): wpe.grad[:T] += column sums
            ops.colsum_(dx.view(B, T * C), ctx.sink_wpe.grad.view(-1)[:T * C])
            ops.sink_done(ctx.sink_wpe)
        elif ctx.needs_input_grad[3]:
            dwpe = torch.zeros((wpe_rows, C), device=dx.device, dtype=torch.float32)   # rows >= T stay zero
            ops.colsum_(dx.view(B, T * C), dwpe.view(-1)[:T * C])
        return None, dprompt, dwte, dwpe, None, None, None


class NormalizeGradientsFn(torch.autograd.Function):
    """reference models/functions.py:4-27: identity forward; backward g / (||g||_2 + 1e-6) over the whole tensor."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        acc = torch.empty(1, device=g.device, dtype=torch.float64)
        call("i2t_gradnorm_scale", ptr(g), ptr(out), ptr(acc), g.numel(), dt(g), stream())
        return out


class LmLossFn(torch.autograd.Function):
    """training/wrapper.py:120-151 (+ get_weights :80-96).  The gradient w.r.t. the logits is produced by the same
    pass over the vocabulary that computes the loss."""

    @staticmethod
    def forward(ctx, logits, teacher_logits, labels, temperature, alpha, weight_fn, eos_weight, eos_id, ignore_index,
                grad_prescale=1.0):
        """grad_prescale != 1: the caller promises to back-propagate exactly `loss * grad_prescale`; the factor is folded
        into the dlogits this pass writes and backward() skips the rescaling pass over the (B,T,V) gradient."""
        B, T_logits, V = logits.shape

        def rows_ok(t):   # (B,T,V) view of a (B*T, pitch) buffer: what the LM head produces (pitch = V rounded up to 8)
            return t.stride(2) == 1 and t.stride(1) >= V and t.stride(0) == T_logits * t.stride(1)
        if not rows_ok(logits):
            logits = logits.contiguous()
        ld = logits.stride(1)
        Tl = min(T_logits, labels.shape[1])
        labels = labels.contiguous()
        need = ctx.needs_input_grad[0]
        weights = torch.empty(B * Tl, device=logits.device, dtype=torch.float32)
        rows = torch.empty(B * Tl, device=logits.device, dtype=torch.float32)
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        dbuf = None
        if need:    # same row pitch as the logits so the LM-head backward GEMMs read it in place
            alloc = torch.zeros if Tl < T_logits else torch.empty
            dbuf = alloc((B, T_logits, ld), device=logits.device, dtype=logits.dtype)
        ld_t = 0
        if teacher_logits is not None:
            teacher_logits = teacher_logits.to(logits.dtype)
            if not rows_ok(teacher_logits):
                teacher_logits = teacher_logits.contiguous()
            ld_t = teacher_logits.stride(1)
        call("i2t_lm_loss", ptr(logits), ptr(teacher_logits), ptr(labels), ptr(weights), ptr(rows), ptr(loss), ptr(dbuf),
             B, T_logits, Tl, V, labels.shape[1], float(temperature), float(alpha if alpha is not None else 0.0),
             int(weight_fn == "inverse_sqrt_position"), int(eos_weight is not None),
             float(eos_weight if eos_weight is not None else 0.0), int(eos_id), int(ignore_index), ld, ld_t,
             float(grad_prescale), dt(logits), stream())
        if need:
            ctx.save_for_backward(dbuf)
            ctx.V = V
            ctx.prescaled = float(grad_prescale) != 1.0
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (dbuf,) = ctx.saved_tensors
        if not ctx.prescaled:
            scale = gloss.reshape(1).to(torch.float32).contiguous()
            call("i2t_scale_inplace", ptr(dbuf), ptr(scale), dbuf.numel(), dt(dbuf), stream())
        return dbuf[..., :ctx.V], None, None, None, None, None, None, None, None, None


class L2NormFn(torch.autograd.Function):
    """F.normalize(x, p=2, dim=-1): reference models/encoder.py:118-119."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        cols = x.shape[-1]
        y = torch.empty_like(x)
        call("i2t_l2norm_fwd", ptr(x), ptr(y), x.numel() // cols, cols, 1e-12, stream())
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        cols = x.shape[-1]
        dx = torch.empty_like(x)
        call("i2t_l2norm_bwd", ptr(x), ptr(dy), ptr(dx), x.numel() // cols, cols, 1e-12, stream())
        return dx


class Conv1DFn(torch.autograd.Function):
    """HF GPT-2 `Conv1D`: y = act(x W + b) + residual with W stored (in, out) -- transformers/pytorch_utils.py Conv1D as
    used by reference models/decoder.py:285-361.  The weight keeps its checkpoint layout; it is the MN-major B operand of
    the forward GEMM and the K-major B operand of the data-gradient GEMM."""

    @staticmethod
    def forward(ctx, x2d, w, w_c, bias, residual, act, out_dtype, drop=None):
        need = any(ctx.needs_input_grad)
        bs = isinstance(w, torch.nn.Parameter)                # (see LinearFn: weights are stable between optimiser steps)
        if drop is not None:                                  # HF resid_pdrop before the residual add
            assert act == ops.ACT_NONE and out_dtype == torch.float32
            z = None
            y = ops.dropout_add(ops.gemm(x2d, w_c, bias=bias, out_dtype=x2d.dtype, b_kmajor=False, b_stable=bs), residual, drop)
        elif need and act != ops.ACT_NONE:
            z = ops.gemm(x2d, w_c, bias=bias, out_dtype=x2d.dtype, b_kmajor=False, b_stable=bs)
            y = torch.empty(z.shape, device=z.device, dtype=out_dtype)
            call("i2t_act_fwd", ptr(z), ptr(y), z.numel(), act, dt(z), dt(y), stream())
            assert residual is None
        else:
            z = None
            y = ops.gemm(x2d, w_c, bias=bias, residual=residual, act=act, out_dtype=out_dtype, b_kmajor=False, b_stable=bs)
        if need:
            ctx.save_for_backward(x2d, w_c, z)
            ctx.act = act
            ctx.has_bias = bias is not None
            ctx.has_res = residual is not None
            ctx.res_dtype = residual.dtype if residual is not None else None
            ctx.drop = drop
            ctx.b_stable = bs
            ctx.sink_w, ctx.sink_b = _sinks_of(ctx, w, bias, None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, w_c, z = ctx.saved_tensors
        dy = dy.contiguous()
        dres = dy.to(ctx.res_dtype) if (ctx.has_res and ctx.needs_input_grad[4]) else None
        g = _grad_in(dy, x2d.dtype, ctx.drop, None)
        if z is not None:
            dz = torch.empty_like(z)
            call("i2t_act_bwd", ptr(z), ptr(g), ptr(dz), z.numel(), ctx.act, dt(z), dt(g), stream())
            g = dz
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(g, w_c, out_dtype=x2d.dtype, a_kmajor=True, b_kmajor=True, b_stable=ctx.b_stable)   # dX = dY W^T, W (K,N) is [n'][k']
        if ctx.needs_input_grad[1]:
            M, N = g.shape
            K = x2d.shape[1]
            if ctx.sink_w is not None:                                                        # dW (K,N) = X^T dY
                pw = ctx.sink_w[0]
                ops.gemm(x2d, g, out=pw.grad, accumulate=True, a_kmajor=False, b_kmajor=False, M=K, N=N, K=M, lda=x2d.stride(0),
                         ldb=g.stride(0), ldc=N)
                ops.sink_done(pw)
            else:
                dw = torch.empty((K, N), device=g.device, dtype=torch.float32)
                ops.gemm(x2d, g, out=dw, a_kmajor=False, b_kmajor=False, M=K, N=N, K=M, lda=x2d.stride(0), ldb=g.stride(0), ldc=N)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = _bias_grad(ctx, g)
        return dx, dw, None, db, dres, None, None, None


def conv1d(x2d, w, w_c, bias=None, residual=None, act=ops.ACT_NONE, out_dtype=torch.float32, drop=None):
    return Conv1DFn.apply(x2d, w, w_c, bias, residual, act, out_dtype, drop)


class PeerLookupFn(torch.autograd.Function):
    """Everything of PeerLookup.forward (reference models/layers.py:73-109) that is not a dense projection: the two top-k
    selections, the top-k of their sums, softmax, expert gather, the dot with the key projection, GELU and the weighted sum
    of the output experts (csrc/peer.cu).  ql, qr: (M, H, U); key: (M, H, D); emb_in (E, D); emb_out (E, O) -> (M, O)."""

    @staticmethod
    def forward(ctx, ql, qr, key, emb_in, emb_out, topk):
        M, H, U = ql.shape
        D, O = key.shape[-1], emb_out.shape[1]
        ql, qr, key = ql.contiguous().float(), qr.contiguous().float(), key.contiguous().float()
        dev = ql.device
        out = torch.empty((M, O), device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)
        f32 = dict(device=dev, dtype=torch.float32)
        idx, lpos, rpos = (torch.empty((M, H, topk), **i32) for _ in range(3))
        score, dot = (torch.empty((M, H, topk), **f32) for _ in range(2))
        call("i2t_peer_lookup_fwd", ptr(ql), ptr(qr), ptr(key), ptr(emb_in), ptr(emb_out), ptr(out), ptr(idx), ptr(score), ptr(dot),
             ptr(lpos), ptr(rpos), M, H, U, topk, D, O, stream())
        ctx.save_for_backward(key, emb_in, emb_out, idx, score, dot, lpos, rpos)
        ctx.dims = (M, H, U, topk, D, O)
        return out

    @staticmethod
    def backward(ctx, dout):
        key, emb_in, emb_out, idx, score, dot, lpos, rpos = ctx.saved_tensors
        M, H, U, K, D, O = ctx.dims
        dout = dout.contiguous().float()
        dev = dout.device
        dql = torch.empty((M, H, U), device=dev, dtype=torch.float32)
        dqr = torch.empty_like(dql)
        dkey = torch.empty((M, H, D), device=dev, dtype=torch.float32)
        demb_in = torch.zeros_like(emb_in)
        demb_out = torch.zeros_like(emb_out)
        call("i2t_peer_lookup_bwd", ptr(dout), ptr(key), ptr(emb_in), ptr(emb_out), ptr(idx), ptr(score), ptr(dot), ptr(lpos),
             ptr(rpos), ptr(dql), ptr(dqr), ptr(dkey), ptr(demb_in), ptr(demb_out), M, H, U, K, D, O, stream())
        return dql, dqr, dkey, demb_in, demb_out, None


class ContrastiveLossFn(torch.autograd.Function):
    """training/wrapper.py:98-118 after the similarity GEMM: masked, weighted cross entropy of pred (R, R) against the diagonal."""

    @staticmethod
    def forward(ctx, pred, labels, L, temperature, weight_fn, eos_weight, eos_id, ignore_index):
        pred = pred.contiguous().float()
        R = pred.shape[0]
        B = R // L
        labels = labels.contiguous()
        dev = pred.device
        weights = torch.empty(R, device=dev, dtype=torch.float32)
        rows = torch.empty(R, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        dpred = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        call("i2t_contrastive_loss", ptr(pred), ptr(labels), ptr(weights), ptr(rows), ptr(loss), ptr(dpred), B, L, labels.shape[1],
             float(temperature), int(weight_fn == "inverse_sqrt_position"), int(eos_weight is not None),
             float(eos_weight if eos_weight is not None else 0.0), int(eos_id), int(ignore_index), 1.0, stream())
        if dpred is not None:
            ctx.save_for_backward(dpred)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (dpred,) = ctx.saved_tensors
        scale = gloss.reshape(1).to(torch.float32).contiguous()
        call("i2t_scale_inplace", ptr(dpred), ptr(scale), dpred.numel(), dt(dpred), stream())
        return dpred, None, None, None, None, None, None, None
