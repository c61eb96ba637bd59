"""Full-sequence forward of the encoder and the decoder (training and the ``forward()`` API).

Restates WHAT the reference computes (citations per function); HOW is B200-specific: packed QKV buffers consumed in
place by the fused attention kernels, masks as closed forms, residual adds / bias / GELU folded into GEMM epilogues,
the LayerNorm output written directly in the GEMM operand dtype.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import call
from .autograd_ops import AttnFn, EmbedFn, LayerNormFn, LayerNormSkipFn, NormalizeGradientsFn, XAttnFn, linear
from .model_spec import layer_has_cross_attn
from .ops import ptr, stream


def _ln(W, key_w, key_b, x, eps, out_dtype):
    return LayerNormFn.apply(x, W[key_w], W.get(key_b), eps, out_dtype)


def _ln_skip(W, key_w, key_b, x, eps, out_dtype):
    """(LayerNorm(x), x): use the returned x for the block's skip connection (one backward pass for both, LayerNormSkipFn)."""
    if torch.is_grad_enabled() and x.requires_grad:
        return LayerNormSkipFn.apply(x, W[key_w], W.get(key_b), eps, out_dtype)
    return LayerNormFn.apply(x, W[key_w], W.get(key_b), eps, out_dtype), x


def _lin(W, key_w, key_b, x2d, residual=None, act=ops.ACT_NONE, out_dtype=torch.float32, rows: Optional[slice] = None,
         drop=None, tok_drop=None):
    w = W[key_w]
    wc = W.c(key_w)
    b = W.get(key_b) if key_b else None
    sink = None
    if rows is not None:
        sink = (w, rows, b)          # ops.grad_sinks(): the gradients of the row views go straight into the packed .grad
        w, wc = w[rows], wc[rows]
        b = b[rows] if b is not None else None
    return linear(x2d, w, wc, b, residual, act, out_dtype, drop=drop, tok_drop=tok_drop, sink=sink)


def _site(drop, p):
    """Next dropout site of this forward pass (None when not training or p == 0)."""
    return drop.site(p) if drop is not None else None


# ------------------------------------------------------------------------------------------------------------
# encoder: reference models/encoder.py:108-119 over torchvision's VisionTransformer.forward (:268-305)
# ------------------------------------------------------------------------------------------------------------
def vit_trunk(W, spec, images: torch.Tensor, cd: torch.dtype, pre: str) -> torch.Tensor:
    p, d, H = spec["vit_patch"], spec["vit_dim"], spec["vit_heads"]
    B = images.shape[0]
    images = images.contiguous().float()
    patches = ops.patch_im2col(images, p, cd)                                   # (B*np, 3*p*p)
    conv_w = W[pre + "conv_proj.weight"]
    po = linear(patches, conv_w.view(d, -1), W.c(pre + "conv_proj.weight").view(d, -1), W[pre + "conv_proj.bias"],
                None, ops.ACT_NONE, torch.float32)                             # (B*np, d)
    x = _AssembleFn.apply(po, W[pre + "class_token"], W[pre + "encoder.pos_embedding"], B)   # (B, np+1, d) fp32
    T = x.shape[1]
    for i in range(spec["vit_layers"]):
        lp = f"{pre}encoder.layers.encoder_layer_{i}."
        y, x = _ln_skip(W, lp + "ln_1.weight", lp + "ln_1.bias", x, 1e-6, cd)
        qkv = _lin(W, lp + "self_attention.in_proj_weight", lp + "self_attention.in_proj_bias", y.view(B * T, d),
                   out_dtype=cd)
        a = AttnFn.apply(qkv, B, T, H, ops.MASK_NONE, 0)
        x = _lin(W, lp + "self_attention.out_proj.weight", lp + "self_attention.out_proj.bias", a,
                 residual=x.view(B * T, d)).view(B, T, d)
        y, x = _ln_skip(W, lp + "ln_2.weight", lp + "ln_2.bias", x, 1e-6, cd)
        h = _lin(W, lp + "mlp.0.weight", lp + "mlp.0.bias", y.view(B * T, d), act=ops.ACT_GELU_ERF, out_dtype=cd)
        x = _lin(W, lp + "mlp.3.weight", lp + "mlp.3.bias", h, residual=x.view(B * T, d)).view(B, T, d)
    # final LayerNorm only on token 0 (the reference normalises all tokens and then keeps x[:, 0])
    cls_tok = x[:, 0, :]
    if torch.is_grad_enabled() and x.requires_grad:
        return LayerNormFn.apply(cls_tok.contiguous(), W[pre + "encoder.ln.weight"], W[pre + "encoder.ln.bias"], 1e-6,
                                 torch.float32)
    return ops.layernorm(x, W[pre + "encoder.ln.weight"], W[pre + "encoder.ln.bias"], 1e-6, torch.float32, rows=B,
                         row_stride=T * d)


class _AssembleFn(torch.autograd.Function):
    """prepend class token, add pos_embedding: torchvision vision_transformer.py:294-296,156."""

    @staticmethod
    def forward(ctx, patch_out, cls, pos, B):
        return ops.vit_assemble(patch_out, cls, pos, B)

    @staticmethod
    def backward(ctx, dx):
        B, T, C = dx.shape
        dx = dx.contiguous()
        dpatch = dx[:, 1:, :].reshape(B * (T - 1), C)
        dcls = dpos = None
        if ctx.needs_input_grad[2]:
            acc = torch.zeros(T * C, device=dx.device, dtype=torch.float32)
            ops.colsum_(dx.view(B, T * C), acc)
            dpos = acc.view(1, T, C)
            if ctx.needs_input_grad[1]:
                dcls = dpos[:, :1, :].clone()
        return dpatch, dcls, dpos, None


class _LshTailFn(torch.autograd.Function):
    """reference models/layers.py:139-144,211-219 + models/encoder.py:116-117 (forward); the backward is the
    EmbeddingBag(mean) scatter: d emb[idx[b,s,r,p], :] += d out[b,s,:] / n_proj (the hashing itself has no gradient)."""

    @staticmethod
    def forward(ctx, feat, tables, num_bins, n_cls, n_proj, E, *emb_weights):
        B, D = feat.shape
        n_res = len(num_bins)
        out = torch.empty((B, n_cls, E), device=feat.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        idx = torch.empty((B, n_cls, n_res, n_proj), device=feat.device, dtype=torch.int32) if need else None
        call("i2t_lsh_tail", ptr(feat), ptr(tables["proj"]), ptr(tables["grid"]), ptr(tables["emb"]), ptr(tables["nb"]),
             ptr(out), ptr(idx), B, D, n_cls, n_res, n_proj, E, stream())
        if need:
            ctx.save_for_backward(idx)
            ctx.meta = (n_cls, n_res, n_proj, E, [w.shape for w in emb_weights])
            # ops.grad_sinks(): all tables or none accumulate into their .grad (one kernel either way)
            ctx.sinks = None
            if all(ops.sink_use(w, count=False) for w in emb_weights):
                ctx.sinks = [w for w in emb_weights if ops.sink_use(w)]
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes
        (idx,) = ctx.saved_tensors
        n_cls, n_res, n_proj, E, shapes = ctx.meta
        dout = dout.contiguous().float()
        B = dout.shape[0]
        if ctx.sinks is not None:
            targets, grads = [w.grad for w in ctx.sinks], [None] * len(shapes)
        else:                      # one zeroed buffer, one view per table
            flat = torch.zeros(sum(s[0] * s[1] for s in shapes), device=dout.device, dtype=torch.float32)
            targets, off = [], 0
            for s in shapes:
                targets.append(flat[off:off + s[0] * s[1]].view(s))
                off += s[0] * s[1]
            grads = targets
        host = (ctypes.c_void_p * len(targets))(*[t.data_ptr() for t in targets])
        call("i2t_lsh_tail_bwd", ptr(dout), ptr(idx), ctypes.addressof(host), B, n_cls, n_res, n_proj, E, stream())
        if ctx.sinks is not None:
            for w in ctx.sinks:
                ops.sink_done(w)
        return (None, None, None, None, None, None, *grads)


def posbias_tail(W, spec, feat: torch.Tensor, cd: torch.dtype, pre: str) -> torch.Tensor:
    """F.normalize -> one MLP (768 -> gate sizes -> out, GELU tanh) per summary slot + residual -> F.normalize:
    reference models/encoder.py:118-119 + models/layers.py:617-638 (AdvancedPositionalBiasMLP) + :222-255 (MLP)."""
    from .autograd_ops import L2NormFn
    x = L2NormFn.apply(feat)                                   # (B, 768) fp32
    xc = x if cd == torch.float32 else x.to(cd)
    outs = []
    n_lin = len(spec["gate_sizes"]) + 1
    for s in range(spec["n_cls"]):
        kp = f"{pre}{s}."
        if (kp + "residual_connector.weight") in W:
            res = _lin(W, kp + "residual_connector.weight", kp + "residual_connector.bias", xc)
        else:
            res = x
        h = xc
        for j in range(n_lin):
            last = j == n_lin - 1
            h = _lin(W, kp + f"model.{2 * j}.weight", kp + f"model.{2 * j}.bias", h, residual=res if last else None,
                     act=ops.ACT_NONE if last else ops.ACT_GELU_TANH, out_dtype=torch.float32 if last else cd)
        outs.append(h)
    return L2NormFn.apply(torch.stack(outs, dim=1))


def peer_tail(W, spec, feat: torch.Tensor, pre: str) -> torch.Tensor:
    """reference models/encoder.py:114-115 + models/layers.py:73-109: per-slot projection of the ViT feature
    (einsum 'bd,des->bse' with peer_proj_wt (768, 768, n_cls): its memory layout is a (in, out) matrix with out = (e, s), i.e. the
    Conv1D GEMM), then PeerLookup: four dense projections (fp32 GEMMs: the expert selection is discrete, see csrc/peer.cu) and the
    fused selection / gather / weighting kernel."""
    from .autograd_ops import PeerLookupFn, conv1d
    B, d = feat.shape
    S, nh, qd, K = spec["n_cls"], spec["peer_nhead"], spec["peer_query_dim"], spec["peer_topk"]
    f32 = torch.float32
    feat = feat.float()
    pw = W[pre + "peer_proj_wt"]
    inp = conv1d(feat, pw.view(d, d * S), pw.view(d, d * S), None, None, ops.ACT_NONE, f32)        # (B, (e, s))
    inp = inp.view(B, d, S).transpose(1, 2).contiguous().view(B * S, d)                            # (B * S, e)

    def lin(key, x):
        w = W[pre + "peer." + key + ".weight"]
        return linear(x, w, w, None, None, ops.ACT_NONE, f32)

    q = lin("query_linear", inp).view(B * S * nh, qd)
    keyp = lin("key_linear", inp).view(B * S, nh, d)
    residual = lin("residual", inp)
    ql = lin("query_left.linear", q).view(B * S, nh, -1)
    qr = lin("query_right.linear", q).view(B * S, nh, -1)
    out = PeerLookupFn.apply(ql, qr, keyp, W[pre + "peer.emb_in.weight"], W[pre + "peer.emb_out.weight"], K)
    return (out + residual).view(B, S, -1)


def encoder_forward(W, spec, images: torch.Tensor, cd: torch.dtype, train_trunk: bool = False) -> torch.Tensor:
    bridged = "encoder.1.weight" in W
    pre = "encoder.0." if bridged else "encoder."
    if train_trunk:
        feat = vit_trunk(W, spec, images, cd, pre + "model.")
    else:
        with torch.no_grad():                                   # reference models/encoder.py:109-111
            feat = vit_trunk(W, spec, images, cd, pre + "model.")
    if spec["tail"] == "lsh":
        tables = W.m._lsh_tables(pre)
        embs = [W[f"{pre}lsh_emb.{s}.emb.{r}.emb.weight"] for s in range(spec["n_cls"])
                for r in range(len(spec["lsh_num_bins"]))]
        out = _LshTailFn.apply(feat, tables, spec["lsh_num_bins"], spec["n_cls"], spec["lsh_num_proj"],
                               spec["n_embd_out_vit"], *embs)
    elif spec["tail"] == "peer":
        out = peer_tail(W, spec, feat, pre)
    else:
        out = posbias_tail(W, spec, feat, cd, pre + "proj.models.")
    if bridged:
        B, n, E = out.shape
        out = linear(out.view(B * n, E).to(cd), W["encoder.1.weight"], W.c("encoder.1.weight"), None, None, ops.ACT_NONE,
                     torch.float32).view(B, n, -1)
    return out


# ------------------------------------------------------------------------------------------------------------
# decoder: reference models/vision_encoder_decoder.py:84-134 + models/decoder.py:214-256 + models/layers.py:565-614
# ------------------------------------------------------------------------------------------------------------
def _given_embeds(inputs_embeds, limit):
    """(ids stand-in, prompt rows, T): every row of the decoder input comes from the caller's inputs_embeds (+ wpe)."""
    T = min(inputs_embeds.shape[1], limit)
    rows = inputs_embeds[:, :T].contiguous().float()
    return torch.zeros((rows.shape[0], 1), dtype=torch.long, device=rows.device), rows, T


def hf_gpt2_forward(W, spec, ids: torch.Tensor, encoder_output: torch.Tensor, cd: torch.dtype, drop=None, inputs_embeds=None):
    """transformers GPT2LMHeadModel with add_cross_attention=True exactly as the reference drives it
    (models/decoder.py:335-361: attention_mask=None -> plain causal over prompt + text; every block has
    ln_cross_attn + crossattention {q_attn, c_attn -> [k|v], c_proj}); Conv1D weights stay (in, out)."""
    from .autograd_ops import conv1d
    C, H = spec["n_embd"], spec["n_head"]
    dp = "decoder.backbone.transformer."
    pd = spec.get("dropout", 0.0)              # GPT2Config embd_pdrop = attn_pdrop = resid_pdrop (training mode only)
    if inputs_embeds is not None:              # HuggingfaceDecoder.forward(inputs_embeds=...), models/decoder.py:335-361
        ids, rows, T = _given_embeds(inputs_embeds, 1024)
        B, n_prompt = rows.shape[0], 0
        x = EmbedFn.apply(ids, rows, W[dp + "wte.weight"], W[dp + "wpe.weight"], T, T, _site(drop, pd))
    else:
        B, S = ids.shape
        n_prompt = spec["n_cls"] if spec["use_soft_prompting"] else 0
        T = min(n_prompt + S, 1024)
        prompt = encoder_output.contiguous().float() if n_prompt else None
        x = EmbedFn.apply(ids.contiguous(), prompt, W[dp + "wte.weight"], W[dp + "wpe.weight"], T, n_prompt, _site(drop, pd))
    cross = spec["use_cross_attn"]
    S_enc = encoder_output.shape[1]
    enc_c = None
    if cross:
        enc_c = encoder_output.reshape(B * S_enc, C)
        enc_c = enc_c if enc_c.dtype == cd else enc_c.to(cd)

    def c1d(wkey, x2d, residual=None, act=ops.ACT_NONE, out_dtype=torch.float32, dsite=None):
        return conv1d(x2d, W[wkey + ".weight"], W.c(wkey + ".weight"), W[wkey + ".bias"], residual, act, out_dtype, dsite)

    for i in range(spec["n_layer"]):
        lp = f"{dp}h.{i}."
        y, x = _ln_skip(W, lp + "ln_1.weight", lp + "ln_1.bias", x, 1e-5, cd)
        qkv = c1d(lp + "attn.c_attn", y.view(B * T, C), out_dtype=cd)
        a = AttnFn.apply(qkv, B, T, H, ops.MASK_CAUSAL, 0, _site(drop, pd))
        x = c1d(lp + "attn.c_proj", a, residual=x.view(B * T, C), dsite=_site(drop, pd)).view(B, T, C)
        if cross:
            y, x = _ln_skip(W, lp + "ln_cross_attn.weight", lp + "ln_cross_attn.bias", x, 1e-5, cd)
            q = c1d(lp + "crossattention.q_attn", y.view(B * T, C), out_dtype=cd)
            kv = c1d(lp + "crossattention.c_attn", enc_c, out_dtype=cd)
            a = XAttnFn.apply(q, kv, B, T, S_enc, H, _site(drop, pd))
            x = c1d(lp + "crossattention.c_proj", a, residual=x.view(B * T, C), dsite=_site(drop, pd)).view(B, T, C)
        y, x = _ln_skip(W, lp + "ln_2.weight", lp + "ln_2.bias", x, 1e-5, cd)
        h = c1d(lp + "mlp.c_fc", y.view(B * T, C), act=ops.ACT_GELU_TANH, out_dtype=cd)
        x = c1d(lp + "mlp.c_proj", h, residual=x.view(B * T, C), dsite=_site(drop, pd)).view(B, T, C)
    hidden = _ln(W, dp + "ln_f.weight", dp + "ln_f.bias", x, 1e-5, torch.float32)
    text = hidden[:, n_prompt:, :].reshape(B * (T - n_prompt), C)
    logits = linear(text, W["decoder.backbone.lm_head.weight"], W.c("decoder.backbone.lm_head.weight"), None, None,
                    ops.ACT_NONE, cd, pad_rows=True).view(B, T - n_prompt, -1)
    return logits, hidden


def decoder_forward(W, spec, ids: torch.Tensor, encoder_output: torch.Tensor, cd: torch.dtype, training: bool = False,
                    drop=None, inputs_embeds=None):
    """`drop` (ops.DropCtx) switches the training-mode dropouts on: transformer.drop (models/decoder.py:236-243), the
    per-token q/k/v masks + SDPA dropout_p + resid_dropout of models/layers.py:454-469, nn.MultiheadAttention's dropout
    (:537-542) and _MLP.dropout (:485).  Site order = call order below (the oracle's mask provider counts the same way)."""
    if spec["decoder"] == "hf_gpt2":
        return hf_gpt2_forward(W, spec, ids, encoder_output, cd, drop, inputs_embeds=inputs_embeds)
    C, H, blk = spec["n_embd"], spec["n_head"], spec["block_size"]
    dp = "decoder.transformer."
    pd, pa = spec.get("dropout", 0.0), spec.get("attn_dropout", 0.0)
    if inputs_embeds is not None:              # TransformerDecoder.forward(inputs_embeds=...), models/decoder.py:231-243
        ids, rows, T = _given_embeds(inputs_embeds, blk)
        B, n_prompt = rows.shape[0], 0
        x = EmbedFn.apply(ids, rows, W[dp + "wte.weight"], W[dp + "wpe.weight"], T, T, _site(drop, pd))
    else:
        B, S = ids.shape
        n_prompt = spec["n_cls"] if spec["use_soft_prompting"] else 0
        T = min(n_prompt + S, blk)
        prompt = encoder_output.contiguous().float() if n_prompt else None
        x = EmbedFn.apply(ids.contiguous(), prompt, W[dp + "wte.weight"], W[dp + "wpe.weight"], T, n_prompt, _site(drop, pd))
    if n_prompt:
        mask_mode = ops.MASK_PROMPT
    else:
        mask_mode = ops.MASK_CAUSAL if spec["is_causal"] else ops.MASK_NONE
    cross = encoder_output if spec["use_cross_attn"] else None
    S_enc = encoder_output.shape[1]
    enc_c = None
    if cross is not None:
        enc_c = encoder_output.reshape(B * S_enc, C)
        enc_c = enc_c if enc_c.dtype == cd else enc_c.to(cd)
    grads = torch.is_grad_enabled()
    for depth in range(spec["n_layer"]):
        lp = f"{dp}h.{depth}."
        y, x = _ln_skip(W, lp + "ln_1.weight", lp + "ln_1.bias", x, 1e-5, cd)
        ts = _site(drop, pa)
        # the backward of the token-level q / k / v masks rides on the attention backward (AttnFn `tok`), not on a pass of its own
        qkv = _lin(W, lp + "attn.c_attn.weight", lp + "attn.c_attn.bias", y.view(B * T, C), out_dtype=cd,
                   tok_drop=(ts, C, 3, True) if ts is not None else None)
        a = AttnFn.apply(qkv, B, T, H, mask_mode, n_prompt, _site(drop, pd), ts)
        x = _lin(W, lp + "attn.c_proj.weight", lp + "attn.c_proj.bias", a, residual=x.view(B * T, C),
                 drop=_site(drop, pd)).view(B, T, C)
        use_cross = cross is not None and (depth % 2 == 0 if spec["skip_alternate_cross_attn"] else True)
        if use_cross:
            if not layer_has_cross_attn(spec, depth):
                raise ValueError("Model not configured for cross attn inputs!!!")
            y, x = _ln_skip(W, lp + "ln_3.weight", lp + "ln_3.bias", x, 1e-5, cd)
            kw, kb = lp + "cross_attn.in_proj_weight", lp + "cross_attn.in_proj_bias"
            q = _lin(W, kw, kb, y.view(B * T, C), out_dtype=cd, rows=slice(0, C))
            kv = _lin(W, kw, kb, enc_c, out_dtype=cd, rows=slice(C, 3 * C))            # raw encoder output, no LN
            a = XAttnFn.apply(q, kv, B, T, S_enc, H, _site(drop, pd))
            x = _lin(W, lp + "cross_attn.out_proj.weight", lp + "cross_attn.out_proj.bias", a,
                     residual=x.view(B * T, C)).view(B, T, C)
        y, x = _ln_skip(W, lp + "ln_2.weight", lp + "ln_2.bias", x, 1e-5, cd)
        h = _lin(W, lp + "mlp.c_fc.weight", lp + "mlp.c_fc.bias", y.view(B * T, C), act=ops.ACT_GELU_TANH, out_dtype=cd)
        x = _lin(W, lp + "mlp.c_proj.weight", lp + "mlp.c_proj.bias", h, residual=x.view(B * T, C),
                 drop=_site(drop, pd)).view(B, T, C)
        if grads and x.requires_grad:
            x = NormalizeGradientsFn.apply(x)                                          # models/layers.py:607-608
    hidden = _ln(W, dp + "ln_f.weight", dp + "ln_f.bias", x, 1e-5, torch.float32)
    # logits only for the text rows (the reference computes all rows, then slices [offset:], :132)
    text = hidden[:, n_prompt:, :].reshape(B * (T - n_prompt), C)
    # logits come out in the compute dtype (bf16 under the reference's autocast) with rows padded to a
    # multiple of 8 elements; the (B, T, V) result is a strided view of that buffer
    logits = linear(text, W["decoder.lm_head.weight"], W.c("decoder.lm_head.weight"), None, None, ops.ACT_NONE,
                    cd, pad_rows=True).view(B, T - n_prompt, -1)
    return logits, hidden
