"""CPU oracle for the image2text encoder-decoder hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32/fp64) restatement of the reference algorithm, written against a
flat ``state_dict`` that uses the reference's checkpoint key names.  It is the checker the
CUDA path is compared with; nothing under ``image2text_b200/`` imports it, and only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may call it.

Parity pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported
from /root/reference under the shims of ``tests/golden/ref_harness.py``) on seeded weights
and inputs and stores its outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function here against those fixtures.  Third-party arithmetic on the path
(torchvision 0.26 ``vit_b_16``, transformers 5.5 ``GPT2LMHeadModel`` with cross attention and
``NoRepeatNGramLogitsProcessor``) is restated from the installed versions and pinned the
same way (the reference itself holds no golden vectors: SURVEY.md section 4).

Every function cites the reference ``file:line`` it follows (paths relative to the
reference root; ``tv:`` = torchvision/models/vision_transformer.py, ``hf:`` =
transformers/models/gpt2/modeling_gpt2.py, ``hfgen:`` = transformers/generation/logits_process.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
NEG_INF = -float("inf")


# ----------------------------------------------------------------------------------------
# spec
# ----------------------------------------------------------------------------------------
def default_spec(**kw) -> dict:
    """Flat description of one model configuration (values of training_configs/local/nano.yaml)."""
    spec = dict(
        # encoder (models/encoder.py:56-127)
        vit_layers=12, vit_heads=12, vit_dim=768, vit_mlp=3072, vit_patch=16, vit_image=224,
        n_cls=8, n_embd_out_vit=768,
        tail="lsh",                 # 'lsh' | 'posbias'
        lsh_num_bins=(4, 8, 20), lsh_num_proj=32,
        # decoder
        decoder="transformer",      # 'transformer' (models/decoder.py:161) | 'hf_gpt2' (models/decoder.py:364)
        n_layer=12, n_head=12, n_embd=768, block_size=256, vocab_size=50257, bias=True,
        ff_mult=4.0, is_causal=True, is_cross_attn=True, skip_alternate_cross_attn=True,
        # top level (configs/models.py:128-135)
        use_cross_attn=True, use_soft_prompting=True, no_repeat_n_grams=(2, 3, 4, 5),
    )
    spec.update(kw)
    return spec


# ----------------------------------------------------------------------------------------
# small ops
# ----------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Optional[Tensor], eps: float) -> Tensor:
    """models/layers.py:357-358 (eps 1e-5); tv:96 (eps 1e-6)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def gelu_tanh(x: Tensor) -> Tensor:
    """models/layers.py:477 nn.GELU(approximate='tanh'); hf 'gelu_new' is the same formula."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x)))


def gelu_erf(x: Tensor) -> Tensor:
    """tv:46 nn.GELU() (exact)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor], pmul: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(hs) + mask) v; a fully masked row yields 0 (torch>=2.5 SDPA).
    pmul: training-mode dropout on the probabilities (dropout_p of models/layers.py:465) as an explicit multiplier
    tensor (0 or 1/(1-p)) so that the caller controls the mask."""
    hs = q.shape[-1]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hs)
    if mask is not None:
        s = s + mask
    m = s.amax(dim=-1, keepdim=True)
    dead = torch.isinf(m) & (m < 0)
    m = torch.where(dead, torch.zeros_like(m), m)
    p = torch.exp(s - m)
    den = p.sum(dim=-1, keepdim=True)
    p = torch.where(dead, torch.zeros_like(p), p / torch.where(dead, torch.ones_like(den), den))
    if pmul is not None:
        p = p * pmul
    return p @ v


def split_heads(x: Tensor, n_head: int) -> Tensor:
    b, t, c = x.shape
    return x.view(b, t, n_head, c // n_head).transpose(1, 2)


def merge_heads(x: Tensor) -> Tensor:
    b, h, t, e = x.shape
    return x.transpose(1, 2).reshape(b, t, h * e)


def packed_mha(sd: Dict[str, Tensor], prefix: str, query: Tensor, kv: Tensor, n_head: int,
               pmul: Optional[Tensor] = None) -> Tensor:
    """torch.nn.MultiheadAttention(batch_first=True), packed in_proj rows [q;k;v], no mask
    (models/layers.py:537-542,600-605; tv:103,113); pmul = its training-mode dropout on the probabilities."""
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    c = query.shape[-1]
    q = F.linear(query, w[:c], b[:c])
    k = F.linear(kv, w[c:2 * c], b[c:2 * c])
    v = F.linear(kv, w[2 * c:], b[2 * c:])
    y = sdpa(split_heads(q, n_head), split_heads(k, n_head), split_heads(v, n_head), None, pmul)
    return F.linear(merge_heads(y), sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


# ----------------------------------------------------------------------------------------
# encoder: torchvision ViT-B/16 trunk + tails  (models/encoder.py:108-119)
# ----------------------------------------------------------------------------------------
def vit_trunk(sd: Dict[str, Tensor], spec: dict, images: Tensor, prefix: str = "encoder.model.") -> Tensor:
    """tv:268-305 (_process_input, forward), tv:154-157 (Encoder.forward), tv:110-119 (EncoderBlock)."""
    p, d, h = spec["vit_patch"], spec["vit_dim"], spec["vit_heads"]
    n = images.shape[0]
    x = F.conv2d(images, sd[prefix + "conv_proj.weight"], sd[prefix + "conv_proj.bias"], stride=p)
    x = x.reshape(n, d, -1).permute(0, 2, 1)                       # (n, 196, d)
    x = torch.cat([sd[prefix + "class_token"].expand(n, -1, -1), x], dim=1)
    x = x + sd[prefix + "encoder.pos_embedding"]
    for i in range(spec["vit_layers"]):
        lp = f"{prefix}encoder.layers.encoder_layer_{i}."
        y = layer_norm(x, sd[lp + "ln_1.weight"], sd[lp + "ln_1.bias"], 1e-6)
        x = x + packed_mha(sd, lp + "self_attention.", y, y, h)
        y = layer_norm(x, sd[lp + "ln_2.weight"], sd[lp + "ln_2.bias"], 1e-6)
        y = gelu_erf(F.linear(y, sd[lp + "mlp.0.weight"], sd[lp + "mlp.0.bias"]))
        x = x + F.linear(y, sd[lp + "mlp.3.weight"], sd[lp + "mlp.3.bias"])
    x = layer_norm(x, sd[prefix + "encoder.ln.weight"], sd[prefix + "encoder.ln.bias"], 1e-6)
    return x[:, 0]


def lsh_bucket_indices(feat: Tensor, projection_mat: Tensor, grid: Tensor, num_bins: int) -> Tensor:
    """models/layers.py:139-143: normalize -> project -> bucketize -> + (num_bins+1)*proj_index.
    Returns int64 (B, n_proj) row indices into the EmbeddingBag table."""
    z = F.normalize(feat, p=2.0, dim=-1) @ projection_mat
    bucket = torch.bucketize(z, grid)                                # count of grid points < z (right=False)
    n_proj = projection_mat.shape[1]
    return bucket + (num_bins + 1) * torch.arange(n_proj, dtype=torch.long, device=feat.device)


def lsh_tail(sd: Dict[str, Tensor], spec: dict, feat: Tensor, prefix: str = "encoder.lsh_emb.") -> Tensor:
    """models/encoder.py:116-117 + models/layers.py:211-219 (sum over resolutions) + :139-144
    (EmbeddingBag mode='mean' over the n_proj rows)."""
    slots = []
    for s in range(spec["n_cls"]):
        acc = None
        for r, nb in enumerate(spec["lsh_num_bins"]):
            kp = f"{prefix}{s}.emb.{r}."
            idx = lsh_bucket_indices(feat, sd[kp + "projection_mat"], sd[kp + "grid"], nb)
            e = sd[kp + "emb.weight"][idx].mean(dim=1)               # (B, emb_dim)
            acc = e if acc is None else acc + e
        slots.append(acc)
    return torch.stack(slots, dim=1)


def posbias_tail(sd: Dict[str, Tensor], spec: dict, feat: Tensor, prefix: str = "encoder.proj.models.") -> Tensor:
    """models/encoder.py:118-119 + models/layers.py:637-638 (one MLP per slot) + :252-255
    (residual; identity connector when in==out, Linear otherwise)."""
    x = F.normalize(feat, p=2.0, dim=-1)
    outs = []
    for s in range(spec["n_cls"]):
        kp = f"{prefix}{s}."
        h = x
        j = 0
        while (kp + f"model.{j}.weight") in sd:
            h = F.linear(h, sd[kp + f"model.{j}.weight"], sd[kp + f"model.{j}.bias"])
            if (kp + f"model.{j + 2}.weight") in sd:
                h = gelu_tanh(h)
            j += 2
        if (kp + "residual_connector.weight") in sd:
            res = F.linear(x, sd[kp + "residual_connector.weight"], sd[kp + "residual_connector.bias"])
        else:
            res = x
        outs.append(h + res)
    return F.normalize(torch.stack(outs, dim=1), p=2.0, dim=-1)


def peer_lookup(sd: Dict[str, Tensor], spec: dict, inp: Tensor, prefix: str = "encoder.peer.") -> Tensor:
    """models/layers.py:73-109 (PeerLookup.forward) with the query units of :21-34.  inp (B, S, 768) -> (B, S, out).
    Literal restatement, including `left * topk + right` at :93-96 (the expert id is NOT left * sqrt(num_units) + right:
    only the first topk * sqrt(num_units) rows of the two tables are ever addressed)."""
    nh, qd, topk = spec["peer_nhead"], spec["peer_query_dim"], spec["peer_topk"]
    bs, seq, fin = inp.shape
    x = F.linear(inp, sd[prefix + "query_linear.weight"]).view(bs, seq, nh, qd)
    inp_proj = F.linear(inp, sd[prefix + "key_linear.weight"]).view(bs, seq, nh, fin)
    residual = F.linear(inp, sd[prefix + "residual.weight"])
    left = torch.topk(F.linear(x, sd[prefix + "query_left.linear.weight"]), k=topk, dim=-1)
    right = torch.topk(F.linear(x, sd[prefix + "query_right.linear.weight"]), k=topk, dim=-1)
    cross = (left.values.unsqueeze(-1) + right.values.unsqueeze(-2)).view(bs, seq, nh, topk * topk)
    y = torch.topk(cross, k=topk, dim=-1)
    scores = F.softmax(y.values, dim=-1)
    li = left.indices.gather(-1, y.indices // topk)
    ri = right.indices.gather(-1, y.indices % topk)
    final = li * topk + ri                                              # (B, S, H, K)
    in_dot = torch.einsum("bshkd,bshd->bshk", sd[prefix + "emb_in.weight"][final], inp_proj)
    weight = scores * gelu_tanh(in_dot)
    return torch.einsum("bshk,bshkd->bsd", weight, sd[prefix + "emb_out.weight"][final]) + residual


def peer_tail(sd: Dict[str, Tensor], spec: dict, feat: Tensor, pre: str = "encoder.") -> Tensor:
    """models/encoder.py:114-115: per-slot projection of the ViT feature, then the expert lookup."""
    return peer_lookup(sd, spec, torch.einsum("bd,des->bse", feat, sd[pre + "peer_proj_wt"]), prefix=pre + "peer.")


def encoder_forward(sd: Dict[str, Tensor], spec: dict, images: Tensor) -> Tensor:
    """models/encoder.py:108-119, plus the bridging Linear of models/vision_encoder_decoder.py:33-37
    (state-dict keys then move under ``encoder.0.`` / ``encoder.1.``)."""
    bridged = "encoder.1.weight" in sd
    pre = "encoder.0." if bridged else "encoder."
    feat = vit_trunk(sd, spec, images, prefix=pre + "model.")
    if spec["tail"] == "lsh":
        out = lsh_tail(sd, spec, feat, prefix=pre + "lsh_emb.")
    elif spec["tail"] == "posbias":
        out = posbias_tail(sd, spec, feat, prefix=pre + "proj.models.")
    elif spec["tail"] == "peer":
        out = peer_tail(sd, spec, feat, pre=pre)
    else:
        raise ValueError(spec["tail"])
    if bridged:
        out = F.linear(out, sd["encoder.1.weight"])
    return out


# ----------------------------------------------------------------------------------------
# decoder: TransformerDecoder (models/decoder.py:214-256) and blocks (models/layers.py:565-614)
# ----------------------------------------------------------------------------------------
class _NormalizeGradients(torch.autograd.Function):
    """models/functions.py:4-27: identity forward, g / (||g||_2 + 1e-6) backward (whole tensor)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g / (torch.norm(g) + 1e-6)


def layer_has_cross_attn(spec: dict, depth: int) -> bool:
    """models/utils.py:39-43 (odd layers lose cross-attn when skip_alternate_cross_attn)."""
    if not spec["is_cross_attn"]:
        return False
    return not (spec["skip_alternate_cross_attn"] and depth % 2 == 1)


def _mul(x: Tensor, m: Optional[Tensor]) -> Tensor:
    return x if m is None else x * m


def transformer_block(sd, spec, lp: str, x: Tensor, cross: Optional[Tensor], attn_mask: Optional[Tensor],
                      has_cross: bool, normalize_grads: bool, drop=None) -> Tensor:
    """models/layers.py:565-614 dense path; attention models/layers.py:447-470; MLP :481-486.
    drop (training mode): a provider of explicit dropout multipliers, asked in call order -- tokens() for the (B,1,T,1)
    q/k/v masks of :454-461, attn() for SDPA's dropout_p :465, elem() for resid_dropout :469, attn() for
    nn.MultiheadAttention's dropout :537-542, elem() for _MLP.dropout :485."""
    nh = spec["n_head"]
    pd, pa = spec.get("dropout", 0.0), spec.get("attn_dropout", 0.0)
    bs, tl = x.shape[0], x.shape[1]
    b_ = (lambda k: sd.get(k)) if spec["bias"] else (lambda k: None)
    if spec["is_causal"]:
        L = x.shape[-2]
        causal = torch.ones((L, L), dtype=torch.bool, device=x.device).tril()
        causal_f = torch.zeros((L, L), dtype=x.dtype, device=x.device).masked_fill(~causal, NEG_INF)[None, None]
        attn_mask = causal_f if attn_mask is None else attn_mask + causal_f
    y = layer_norm(x, sd[lp + "ln_1.weight"], b_(lp + "ln_1.bias"), 1e-5)
    qkv = F.linear(y, sd[lp + "attn.c_attn.weight"], b_(lp + "attn.c_attn.bias"))
    q, k, v = qkv.split(spec["n_embd"], dim=2)
    pmul = None
    if drop is not None:
        tok = drop.tokens(pa, bs * tl)
        if tok is not None:
            tok = tok.view(bs, tl, 3)
            q, k, v = q * tok[..., 0:1], k * tok[..., 1:2], v * tok[..., 2:3]
        pmul = drop.attn(pd, bs, nh, tl, tl)
    y = sdpa(split_heads(q, nh), split_heads(k, nh), split_heads(v, nh), attn_mask, pmul)
    y = F.linear(merge_heads(y), sd[lp + "attn.c_proj.weight"], b_(lp + "attn.c_proj.bias"))
    x = x + _mul(y, drop.elem(pd, y.shape) if drop is not None else None)
    if cross is not None:
        if not has_cross:
            raise ValueError("Model not configured for cross attn inputs!!!")
        y = layer_norm(x, sd[lp + "ln_3.weight"], b_(lp + "ln_3.bias"), 1e-5)
        pmul = drop.attn(pd, bs, nh, tl, cross.shape[1]) if drop is not None else None
        x = x + packed_mha(sd, lp + "cross_attn.", y, cross, nh, pmul)
    y = layer_norm(x, sd[lp + "ln_2.weight"], b_(lp + "ln_2.bias"), 1e-5)
    y = gelu_tanh(F.linear(y, sd[lp + "mlp.c_fc.weight"], b_(lp + "mlp.c_fc.bias")))
    y = F.linear(y, sd[lp + "mlp.c_proj.weight"], b_(lp + "mlp.c_proj.bias"))
    x = x + _mul(y, drop.elem(pd, y.shape) if drop is not None else None)
    if normalize_grads:
        x = _NormalizeGradients.apply(x)
    return x


def transformer_decoder_forward(sd, spec, idx=None, inputs_embeds=None, cross_attn_embeds=None, attn_msk=None,
                                prefix: str = "decoder.", normalize_grads: bool = True, drop=None):
    """models/decoder.py:214-256; drop=None: eval / p=0, else explicit multipliers (transformer.drop :236-243 first)."""
    assert (idx is None) != (inputs_embeds is None)
    if inputs_embeds is None:
        inputs_embeds = sd[prefix + "transformer.wte.weight"][idx]
    t = inputs_embeds.shape[1]
    assert t <= spec["block_size"]
    x = inputs_embeds + sd[prefix + "transformer.wpe.weight"][:t]
    if drop is not None:
        x = _mul(x, drop.elem(spec.get("dropout", 0.0), x.shape))
    for depth in range(spec["n_layer"]):
        if spec["skip_alternate_cross_attn"]:
            cross = cross_attn_embeds if depth % 2 == 0 else None
        else:
            cross = cross_attn_embeds
        x = transformer_block(sd, spec, f"{prefix}transformer.h.{depth}.", x, cross, attn_msk,
                              layer_has_cross_attn(spec, depth), normalize_grads, drop)
    b_ = sd.get(prefix + "transformer.ln_f.bias") if spec["bias"] else None
    x = layer_norm(x, sd[prefix + "transformer.ln_f.weight"], b_, 1e-5)
    return F.linear(x, sd[prefix + "lm_head.weight"]), x


def hf_gpt2_forward(sd, spec, idx=None, inputs_embeds=None, cross_attn_embeds=None,
                    prefix: str = "decoder.backbone.", drop=None):
    """transformers GPT2LMHeadModel with add_cross_attention=True as the reference calls it
    (models/decoder.py:335-361: attention_mask=None -> plain causal; hf: GPT2Block.forward,
    GPT2Attention.forward).  Conv1D weights are stored (in, out): y = x @ W + b."""
    nh = spec["n_head"]
    c = spec["n_embd"]
    if inputs_embeds is None:
        inputs_embeds = sd[prefix + "transformer.wte.weight"][idx]
    t = inputs_embeds.shape[1]
    x = inputs_embeds + sd[prefix + "transformer.wpe.weight"][:t]
    pd = spec.get("dropout", 0.0)          # GPT2Config embd_pdrop = attn_pdrop = resid_pdrop (training mode, drop given)
    bs = x.shape[0]
    dm = (lambda shape: drop.elem(pd, shape)) if drop is not None else (lambda shape: None)
    da = (lambda tk: drop.attn(pd, bs, nh, t, tk)) if drop is not None else (lambda tk: None)
    x = _mul(x, dm(x.shape))
    causal = torch.ones((t, t), dtype=torch.bool, device=x.device).tril()
    mask = torch.zeros((t, t), dtype=x.dtype, device=x.device).masked_fill(~causal, NEG_INF)[None, None]
    for i in range(spec["n_layer"]):
        lp = f"{prefix}transformer.h.{i}."
        y = layer_norm(x, sd[lp + "ln_1.weight"], sd[lp + "ln_1.bias"], 1e-5)
        qkv = y @ sd[lp + "attn.c_attn.weight"] + sd[lp + "attn.c_attn.bias"]
        q, k, v = qkv.split(c, dim=2)
        y = merge_heads(sdpa(split_heads(q, nh), split_heads(k, nh), split_heads(v, nh), mask, da(t)))
        y = y @ sd[lp + "attn.c_proj.weight"] + sd[lp + "attn.c_proj.bias"]
        x = x + _mul(y, dm(y.shape))
        if cross_attn_embeds is not None:
            y = layer_norm(x, sd[lp + "ln_cross_attn.weight"], sd[lp + "ln_cross_attn.bias"], 1e-5)
            q = y @ sd[lp + "crossattention.q_attn.weight"] + sd[lp + "crossattention.q_attn.bias"]
            kv = cross_attn_embeds @ sd[lp + "crossattention.c_attn.weight"] + sd[lp + "crossattention.c_attn.bias"]
            k, v = kv.split(c, dim=2)
            y = merge_heads(sdpa(split_heads(q, nh), split_heads(k, nh), split_heads(v, nh), None,
                                 da(cross_attn_embeds.shape[1])))
            y = y @ sd[lp + "crossattention.c_proj.weight"] + sd[lp + "crossattention.c_proj.bias"]
            x = x + _mul(y, dm(y.shape))
        y = layer_norm(x, sd[lp + "ln_2.weight"], sd[lp + "ln_2.bias"], 1e-5)
        y = gelu_tanh(y @ sd[lp + "mlp.c_fc.weight"] + sd[lp + "mlp.c_fc.bias"])
        y = y @ sd[lp + "mlp.c_proj.weight"] + sd[lp + "mlp.c_proj.bias"]
        x = x + _mul(y, dm(y.shape))
    x = layer_norm(x, sd[prefix + "transformer.ln_f.weight"], sd[prefix + "transformer.ln_f.bias"], 1e-5)
    return F.linear(x, sd[prefix + "lm_head.weight"]), x


def decoder_block_size(spec) -> int:
    return 1024 if spec["decoder"] == "hf_gpt2" else spec["block_size"]   # models/decoder.py:374-376


# ----------------------------------------------------------------------------------------
# top level: VisionEncoderDecoder.forward (models/vision_encoder_decoder.py:51-134)
# ----------------------------------------------------------------------------------------
def expand_user_mask(attn_msk: Tensor, bs: int) -> Tensor:
    """models/vision_encoder_decoder.py:61-72 (einops repeat restated with expand)."""
    if attn_msk.dim() == 2:
        s = attn_msk.shape[1]
        if attn_msk.shape[0] == bs:
            return attn_msk[:, None, :, None].expand(bs, 1, s, s)       # 'bs s -> bs h s l' (query rows!)
        return attn_msk[None, None].expand(bs, 1, *attn_msk.shape)      # 's l -> bs h s l'
    if attn_msk.dim() == 3:
        if attn_msk.shape[0] == bs:
            return attn_msk[:, None]                                    # 'bs s l -> bs h s l'
        return attn_msk[None].expand(bs, *attn_msk.shape)               # 'h s l -> bs h s l'
    return attn_msk


def _bool_mask_to_float(attn_msk: Tensor) -> Tensor:
    """models/vision_encoder_decoder.py:101-102 and :117-118, restated LITERALLY.

    ``masked_fill`` runs on the BOOL tensor, so ``-inf`` is cast to ``True``; after ``.float()`` every
    entry is 1.0 and the next line turns every 1 into 0.  The result is all zeros: the caller's
    ``attn_msk`` (padding rows, 2-D / 3-D masks) and the causal AND of :75-82 have NO effect on the
    reference's outputs (pinned by tests/golden/tiny_fwd.npz: logits with a row mask, a random 2-D
    mask and no mask are bit-identical).  Causality comes only from TransformerBlock
    (models/layers.py:581-595) / HF's own causal mask.  DESIGN.md lists this as discrepancy D9.
    """
    m = attn_msk.masked_fill(~attn_msk, NEG_INF).float()
    m[m == 1] = 0
    return m


def ved_forward(sd, spec, images: Optional[Tensor], ids: Tensor, attn_msk: Optional[Tensor] = None,
                encoder_output: Optional[Tensor] = None, normalize_grads: bool = True, drop=None):
    """Returns (encoder_output, logits, hidden_state) like VisionEncoderDecoderModelOutput."""
    if encoder_output is None:
        encoder_output = encoder_forward(sd, spec, images)
    bs = encoder_output.shape[0]
    if attn_msk is not None:
        attn_msk = expand_user_mask(attn_msk, bs)
    L = ids.shape[-1]
    causal = torch.ones((L, L), dtype=torch.bool, device=ids.device).tril()[None, None]
    attn_msk = causal if attn_msk is None else torch.logical_and(attn_msk, causal)
    blk = decoder_block_size(spec)
    dec_prefix = "decoder."
    wte_key = "decoder.backbone.transformer.wte.weight" if spec["decoder"] == "hf_gpt2" \
        else "decoder.transformer.wte.weight"
    if spec["use_soft_prompting"]:
        inputs_embeds = torch.cat((encoder_output, sd[wte_key][ids]), dim=-2)[..., :blk, :].contiguous()
        ncls = encoder_output.shape[1]
        _, h, s, _ = attn_msk.shape
        new = torch.full((bs, h, ncls + s, ncls + s), NEG_INF, dtype=encoder_output.dtype, device=ids.device)
        new[..., :ncls, :] = 0                                           # query rows
        new[..., ncls:, ncls:] = _bool_mask_to_float(attn_msk)
        mask = new[..., :blk, :blk].contiguous()
        dec_ids, offset = None, ncls
    else:
        inputs_embeds, offset, dec_ids = None, 0, ids
        mask = _bool_mask_to_float(attn_msk)
    cross = encoder_output if spec["use_cross_attn"] else None
    if spec["decoder"] == "hf_gpt2":
        logits, hidden = hf_gpt2_forward(sd, spec, idx=dec_ids, inputs_embeds=inputs_embeds, cross_attn_embeds=cross,
                                         drop=drop)
    else:
        logits, hidden = transformer_decoder_forward(sd, spec, idx=dec_ids, inputs_embeds=inputs_embeds,
                                                     cross_attn_embeds=cross, attn_msk=mask, prefix=dec_prefix,
                                                     normalize_grads=normalize_grads, drop=drop)
    return encoder_output, logits[..., offset:, :].contiguous(), hidden


# ----------------------------------------------------------------------------------------
# sampling: VisionEncoderDecoder.generate (models/vision_encoder_decoder.py:136-182)
# ----------------------------------------------------------------------------------------
def banned_ngram_tokens(row: Sequence[int], n: int):
    """hfgen:1012-1076: tokens that would complete an n-gram already present in ``row``."""
    cur_len = len(row)
    if cur_len + 1 < n:
        return []
    key = tuple(row[cur_len + 1 - n:cur_len])
    banned = []
    for i in range(cur_len - n + 1):
        if tuple(row[i:i + n - 1]) == key:
            banned.append(row[i + n - 1])
    return banned


def apply_ngram_ban(ids: Tensor, scores: Tensor, ngrams: Sequence[int]) -> Tensor:
    """hfgen:1127-1135 applied once per configured n (models/vision_encoder_decoder.py:40-43,153)."""
    scores = scores.clone()
    rows = ids.tolist()
    for n in ngrams:
        for i, row in enumerate(rows):
            b = banned_ngram_tokens(row, n)
            if b:
                scores[i, b] = NEG_INF
    return scores


def next_token_probs(logits_last: Tensor, ids: Tensor, spec: dict, temperature: float, top_k: Optional[int]) -> Tensor:
    """models/vision_encoder_decoder.py:152-159: / temperature -> n-gram ban -> top-k threshold
    (ties at the threshold are kept) -> softmax."""
    logits = logits_last / temperature
    logits = apply_ngram_ban(ids, logits, spec["no_repeat_n_grams"])
    if top_k is not None:
        v, _ = torch.topk(logits, min(top_k, logits.shape[-1]), dim=-1)
        logits = logits.masked_fill(logits < v[..., [-1]], NEG_INF)
    return logits.softmax(dim=-1)


def nucleus_filter(probs: Tensor, nucleus_p: float):
    """models/vision_encoder_decoder.py:160-172.  Returns (sorted_probs renormalised, sorted_indices)."""
    sp, si = torch.sort(probs, descending=True, dim=-1)
    cp = torch.cumsum(sp, dim=-1)
    thr = torch.maximum(nucleus_p * torch.ones_like(sp[:, 0]), sp[:, 0]).unsqueeze(1)
    sp = sp.masked_fill(cp > thr, 0.0)
    return sp / sp.sum(dim=-1, keepdim=True), si


@torch.no_grad()
def generate(sd, spec, images, prompt_ids, max_new_tokens=128, temperature=1.0, top_k=None, nucleus_p=None,
             generator: Optional[torch.Generator] = None, return_probs: bool = False):
    """Cache-less loop of the reference: a FULL forward per new token."""
    blk = decoder_block_size(spec) - (spec["n_cls"] if spec["use_soft_prompting"] else 0)
    assert max_new_tokens <= blk - prompt_ids.shape[-1]
    enc = None
    ids = prompt_ids
    all_probs = []
    for _ in range(max_new_tokens):
        cond = ids if ids.shape[-1] <= blk else ids[..., -blk:].contiguous()
        enc, logits, _ = ved_forward(sd, spec, images, cond, encoder_output=enc, normalize_grads=False)
        probs = next_token_probs(logits[..., -1, :], ids, spec, temperature, top_k)
        if return_probs:
            all_probs.append(probs)
        if nucleus_p is not None:
            sp, si = nucleus_filter(probs, nucleus_p)
            nxt = si.gather(-1, torch.multinomial(sp, 1, generator=generator))
        else:
            nxt = torch.multinomial(probs, 1, generator=generator)
        ids = torch.cat((ids, nxt), dim=-1)
    if return_probs:
        return ids, torch.stack(all_probs, dim=1)
    return ids


# ----------------------------------------------------------------------------------------
# training: ModelTrainerWrapper (training/wrapper.py)
# ----------------------------------------------------------------------------------------
def loss_weights(labels: Tensor, weight_fn: str = "constant", eos_token_weight=None, eos_token_id=50256,
                 ignore_index: int = -100) -> Tensor:
    """training/wrapper.py:80-96."""
    if weight_fn == "constant":
        w = torch.ones_like(labels, dtype=torch.float)
    elif weight_fn == "inverse_sqrt_position":
        w = (1.0 / torch.sqrt(torch.arange(1, labels.shape[1] + 1, dtype=torch.float, device=labels.device)))
        w = w.unsqueeze(0).expand(labels.shape[0], -1).clone()
    else:
        raise ValueError(f"unknown weight_fn: {weight_fn}")
    if eos_token_weight is not None:
        w[labels == eos_token_id] = eos_token_weight
    w[labels == ignore_index] = 0.0
    return (w / (1e-3 + w.sum(dim=-1, keepdim=True))) / w.shape[0]


def lm_loss(logits: Tensor, labels: Tensor, logits_moco: Optional[Tensor] = None, temperature: float = 1.0,
            alpha: Optional[float] = None, ignore_index: int = -100, **wkw) -> Tensor:
    """training/wrapper.py:120-151 (weighted CE; distillation soft-CE when a teacher is given)."""
    labels = labels[..., :logits.shape[-2]].contiguous()
    if logits.shape[-2] > labels.shape[-1]:
        logits = logits[..., :labels.shape[-1], :]
        if logits_moco is not None:
            logits_moco = logits_moco[..., :labels.shape[-1], :]
    w = loss_weights(labels, ignore_index=ignore_index, **wkw)
    logp = F.log_softmax(logits / temperature, dim=-1)
    valid = labels != ignore_index
    safe = torch.where(valid, labels, torch.zeros_like(labels))
    picked = logp.gather(-1, safe.unsqueeze(-1)).squeeze(-1) * valid
    if logits_moco is not None:
        soft = F.softmax(logits_moco / temperature, dim=-1)
        per_tok = alpha * (logp * soft).sum(dim=-1) + (1 - alpha) * picked
        return -(per_tok * w).sum()
    return -(picked * w).sum()


def contrastive_loss(sd, hidden_state: Tensor, labels: Tensor, temperature: float = 1.0, ignore_index: int = -100,
                     wte_key: str = "decoder.transformer.wte.weight", **wkw) -> Tensor:
    """training/wrapper.py:98-118: hidden rows vs. the input embeddings of every label of the batch, CE against the diagonal over
    the non-ignored columns, weighted like the LM loss; rows whose own column is masked (infinite CE) count as 0."""
    labels = labels[..., :hidden_state.size(-2)].contiguous()
    if hidden_state.size(-2) > labels.size(-1):
        hidden_state = hidden_state[..., :labels.size(-1), :]
    weights = loss_weights(labels, ignore_index=ignore_index, **wkw)
    keep = labels != ignore_index
    target = sd[wte_key][torch.where(keep, labels, torch.zeros_like(labels))]
    pred = hidden_state.reshape(-1, hidden_state.size(-1)) @ target.reshape(-1, target.size(-1)).T
    pred = torch.where(keep.view(1, -1), pred, torch.full_like(pred, float("-inf")))
    tgt = torch.arange(pred.size(0), device=pred.device)
    losses = F.cross_entropy(pred / temperature, tgt, reduction="none")
    losses = torch.where(losses.isinf(), torch.zeros_like(losses), losses)
    return (losses.view(-1) * weights.view(-1)).sum()


def wrapper_inputs(labels: Tensor, eos_token_id=50256, bos_token_id=50256, ignore_index=-100):
    """training/wrapper.py:154-159,184-196 (no MLM corruption: mask_fraction = 0).
    Returns (decoder input ids with BOS prepended and last dropped, (B,s) row-valid mask)."""
    ids = torch.where(labels != ignore_index, labels, torch.full_like(labels, eos_token_id))
    msk = labels != ignore_index
    bs, sl = ids.shape
    ids = torch.cat((torch.full((bs, 1), bos_token_id, dtype=torch.long, device=ids.device), ids), dim=1)[:, :sl]
    msk = torch.cat((torch.ones((bs, 1), dtype=torch.bool, device=msk.device), msk), dim=1)[:, :sl]
    return ids, msk


def train_step_loss(sd, spec, images, labels, sd_teacher=None, alpha=None, temperature=1.0, **wkw) -> Tensor:
    """training/wrapper.py:153-214 forward + loss (EMA update is a separate call, see ema_update)."""
    ids, msk = wrapper_inputs(labels)
    _, logits, _ = ved_forward(sd, spec, images, ids, attn_msk=msk)
    logits_m = None
    if sd_teacher is not None:
        with torch.no_grad():
            _, logits_m, _ = ved_forward(sd_teacher, spec, images, ids, attn_msk=msk)
    return lm_loss(logits, labels, logits_m, temperature=temperature, alpha=alpha, **wkw)


@torch.no_grad()
def ema_update(params_m: Sequence[Tensor], params: Sequence[Tensor], momentum: float):
    """training/wrapper.py:53-60: p_m = p_m * m + p * (1 - m)."""
    for pm, p in zip(params_m, params):
        pm.copy_(pm * momentum + p * (1.0 - momentum))


@torch.no_grad()
def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float,
               eps: float = 1e-8, weight_decay: float = 0.0):
    """torch.optim.AdamW single-tensor semantics (trainer.py:169-172); ``step`` counts from 1."""
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


@torch.no_grad()
def snradam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float,
                 eps: float = 1e-8, weight_decay: float = 0.0):
    """models/optimizer.py:56-113; ``step`` is the reference's ``iter_`` (starts at 1)."""
    if weight_decay != 0:
        p.mul_(1 - lr * weight_decay)
    if step == 1:
        d = g - m
    else:
        d = g - m * (1.0 / (1 - beta1 ** (step - 1)))
    d = d * d
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).add_(d, alpha=1.0 - beta2)
    p.addcdiv_(m * (1.0 / (1 - beta1 ** step)), (v * (1.0 / (1 - beta2 ** step))).sqrt() + eps, value=-lr)
