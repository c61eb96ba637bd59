"""Flat model description + checkpoint key schema + seeded synthetic weights.

``spec_from_config`` flattens a ``VisionEncoderDecoderConfig`` into the small dict the kernels'
host side (and, independently, the test oracle) work from.  ``state_schema`` lists every
``state_dict`` key the reference model exposes for that configuration with its shape
(reference ``models/decoder.py:169-190``, ``models/encoder.py:56-106``,
``models/layers.py:117-137``) -- this is the checkpoint layout contract.

``synth_state_dict`` fills that schema from a per-key seeded CPU generator.  Hub checkpoints are
unreachable offline, and the reference's own constructors cannot run on the GPU box, so parity
fixtures use these weights on both sides: the unmodified reference loads them with
``load_state_dict`` when the golden vectors are made, and the B200 path loads the same dict.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config_schema import (
    HuggingfaceDecoderConfig,
    MLPConfig,
    PretrainedViTConfig,
    SelfAttentionType,
    TransformerDecoderConfig,
    VisionEncoderDecoderConfig,
)


def spec_from_config(cfg: VisionEncoderDecoderConfig, **overrides) -> dict:
    enc, dec = cfg.vision_encoder_config, cfg.decoder_config
    if not isinstance(enc, PretrainedViTConfig):
        raise NotImplementedError("VisionTransformerEncoder (reference models/encoder.py:130-195) is a "
                                  "'next' row (SURVEY.md 8f-1); only PretrainedViT is built")
    if enc.lora_spec is not None or dec.lora_spec is not None:
        raise NotImplementedError("LoRA needs peft, which is not part of the hot path (SURVEY.md section 2)")
    spec = dict(
        vit_layers=12, vit_heads=12, vit_dim=768, vit_mlp=3072, vit_patch=16, vit_image=224,
        n_cls=enc.n_cls, n_embd_out_vit=enc.n_embd_out_vit,
        refine_base_model=bool(enc.refine_base_model) and enc.lsh_config is None,
        gate_sizes=tuple(enc.gate_sizes) if enc.gate_sizes else (),
        use_cross_attn=cfg.use_cross_attn, use_soft_prompting=cfg.use_soft_prompting,
        no_repeat_n_grams=tuple(cfg.no_repeat_n_grams),
    )
    if enc.peer_config is not None:            # PretrainedViT: PEER wins over LSH (reference models/encoder.py:64-66)
        pc = enc.peer_config
        spec.update(tail="peer", lsh_num_bins=(), lsh_num_proj=0, peer_units_sqrt=pc.num_units_sqrt, peer_topk=pc.topk,
                    peer_nhead=pc.nhead, peer_query_dim=pc.query_dim or 768 // 2,
                    refine_base_model=bool(enc.refine_base_model))
    elif enc.lsh_config is not None:
        if enc.lsh_config.learnable:
            raise NotImplementedError("learnable LSH tail (reference models/layers.py:156-191) is not on the hot path")
        spec.update(tail="lsh", lsh_num_bins=tuple(enc.lsh_config.num_bins), lsh_num_proj=enc.lsh_config.num_proj)
    else:
        spec.update(tail="posbias", lsh_num_bins=(), lsh_num_proj=0)
    if isinstance(dec, TransformerDecoderConfig):
        tc = dec.transformer_config
        ac = tc.attn_config
        if ac.attn_type != SelfAttentionType.MULTI_HEAD or tc.is_sparse_attn or not isinstance(tc.rotator_config, MLPConfig) \
                or dec.use_advanced_pos_emb:
            raise NotImplementedError("multi-query / sparse / MoE / advanced-pos-emb decoder blocks are 'next' rows "
                                      "(SURVEY.md 8f-1)")
        spec.update(decoder="transformer", n_layer=dec.n_layer, n_head=ac.n_head, n_embd=ac.n_embd,
                    block_size=dec.block_size, vocab_size=dec.vocab_size, bias=ac.bias,
                    ff_mult=float(tc.rotator_config.ff_mult), is_causal=tc.is_causal, is_cross_attn=tc.is_cross_attn,
                    skip_alternate_cross_attn=dec.skip_alternate_cross_attn,
                    dropout=ac.dropout, attn_dropout=ac.attn_dropout)
    elif isinstance(dec, HuggingfaceDecoderConfig):
        if not dec.model_str.startswith("gpt2") or dec.load_in_4bit:
            raise NotImplementedError("only the HF GPT-2 layout is built (Falcon path is broken in the reference, "
                                      "SURVEY.md D6; Llama/Qwen need hub weights)")
        dims = {"gpt2": (12, 12, 768), "gpt2-medium": (24, 16, 1024), "gpt2-large": (36, 20, 1280),
                "gpt2-xl": (48, 25, 1600)}[dec.model_str]
        spec.update(decoder="hf_gpt2", n_layer=dims[0], n_head=dims[1], n_embd=dims[2], block_size=1024,
                    vocab_size=dec.vocab_size + dec.extra_tokens, bias=True, ff_mult=4.0, is_causal=True,
                    is_cross_attn=dec.use_cross_attn, skip_alternate_cross_attn=False, dropout=0.1, attn_dropout=0.1)
    else:
        raise ValueError("unknown decoder config")
    spec.update(overrides)
    return spec


def layer_has_cross_attn(spec: dict, depth: int) -> bool:
    """reference models/utils.py:39-43."""
    if not spec["is_cross_attn"]:
        return False
    return not (spec["skip_alternate_cross_attn"] and depth % 2 == 1)


def state_schema(spec: dict) -> "OrderedDict[str, Tuple[Tuple[int, ...], torch.dtype]]":
    """key -> (shape, dtype) for every entry of the reference model's ``state_dict()``."""
    f32, i64 = torch.float32, torch.int64
    out: "OrderedDict[str, Tuple[Tuple[int, ...], torch.dtype]]" = OrderedDict()
    d, mlp, p = spec["vit_dim"], spec["vit_mlp"], spec["vit_patch"]
    seq = (spec["vit_image"] // p) ** 2 + 1
    bridged = spec["n_embd_out_vit"] != spec["n_embd"]
    e = "encoder.0." if bridged else "encoder."
    out[e + "peer_proj_wt"] = ((d, d, spec["n_cls"]) if spec["tail"] == "peer" else (1,), f32)
    out[e + "model.class_token"] = ((1, 1, d), f32)
    out[e + "model.conv_proj.weight"] = ((d, 3, p, p), f32)
    out[e + "model.conv_proj.bias"] = ((d,), f32)
    out[e + "model.encoder.pos_embedding"] = ((1, seq, d), f32)
    for i in range(spec["vit_layers"]):
        lp = f"{e}model.encoder.layers.encoder_layer_{i}."
        out[lp + "ln_1.weight"] = ((d,), f32)
        out[lp + "ln_1.bias"] = ((d,), f32)
        out[lp + "self_attention.in_proj_weight"] = ((3 * d, d), f32)
        out[lp + "self_attention.in_proj_bias"] = ((3 * d,), f32)
        out[lp + "self_attention.out_proj.weight"] = ((d, d), f32)
        out[lp + "self_attention.out_proj.bias"] = ((d,), f32)
        out[lp + "ln_2.weight"] = ((d,), f32)
        out[lp + "ln_2.bias"] = ((d,), f32)
        out[lp + "mlp.0.weight"] = ((mlp, d), f32)
        out[lp + "mlp.0.bias"] = ((mlp,), f32)
        out[lp + "mlp.3.weight"] = ((d, mlp), f32)
        out[lp + "mlp.3.bias"] = ((d,), f32)
    out[e + "model.encoder.ln.weight"] = ((d,), f32)
    out[e + "model.encoder.ln.bias"] = ((d,), f32)
    eo = spec["n_embd_out_vit"]
    if spec["tail"] == "peer":                 # reference models/layers.py:36-71 (registration order of PeerLookup.__init__)
        nh, qd, nq = spec["peer_nhead"], spec["peer_query_dim"], spec["peer_units_sqrt"]
        out[e + "peer.residual.weight"] = ((eo, d), f32)
        out[e + "peer.query_linear.weight"] = ((qd * nh, d), f32)
        out[e + "peer.key_linear.weight"] = ((d * nh, d), f32)
        out[e + "peer.query_left.linear.weight"] = ((nq, qd), f32)
        out[e + "peer.query_right.linear.weight"] = ((nq, qd), f32)
        out[e + "peer.emb_in.weight"] = ((nq * nq, d), f32)
        out[e + "peer.emb_out.weight"] = ((nq * nq, eo), f32)
    elif spec["tail"] == "lsh":
        for s in range(spec["n_cls"]):
            for r, nb in enumerate(spec["lsh_num_bins"]):
                kp = f"{e}lsh_emb.{s}.emb.{r}."
                out[kp + "projection_mat"] = ((d, spec["lsh_num_proj"]), f32)
                out[kp + "grid"] = ((nb,), f32)
                out[kp + "pos_offset"] = ((spec["lsh_num_proj"], 1, 1), i64)
                out[kp + "emb.weight"] = (((nb + 1) * spec["lsh_num_proj"], eo), f32)
    else:
        for s in range(spec["n_cls"]):
            kp = f"{e}proj.models.{s}."
            prev, j = d, 0
            for g in spec["gate_sizes"]:
                out[kp + f"model.{j}.weight"] = ((g, prev), f32)
                out[kp + f"model.{j}.bias"] = ((g,), f32)
                prev, j = g, j + 2
            out[kp + f"model.{j}.weight"] = ((eo, prev), f32)
            out[kp + f"model.{j}.bias"] = ((eo,), f32)
            if eo != d:
                out[kp + "residual_connector.weight"] = ((eo, d), f32)
                out[kp + "residual_connector.bias"] = ((eo,), f32)
    if bridged:
        out["encoder.1.weight"] = ((spec["n_embd"], eo), f32)
    c, v, L = spec["n_embd"], spec["vocab_size"], spec["n_layer"]
    if spec["decoder"] == "transformer":
        ff = int(spec["ff_mult"] * c)
        dp = "decoder.transformer."
        out[dp + "wte.weight"] = ((v, c), f32)
        out[dp + "wpe.weight"] = ((spec["block_size"], c), f32)
        for i in range(L):
            lp = f"{dp}h.{i}."
            names = [("ln_1.weight", (c,)), ("ln_1.bias", (c,)),
                     ("attn.c_attn.weight", (3 * c, c)), ("attn.c_attn.bias", (3 * c,)),
                     ("attn.c_proj.weight", (c, c)), ("attn.c_proj.bias", (c,)),
                     ("ln_2.weight", (c,)), ("ln_2.bias", (c,)),
                     ("mlp.c_fc.weight", (ff, c)), ("mlp.c_fc.bias", (ff,)),
                     ("mlp.c_proj.weight", (c, ff)), ("mlp.c_proj.bias", (c,))]
            if layer_has_cross_attn(spec, i):
                names += [("cross_attn.in_proj_weight", (3 * c, c)), ("cross_attn.in_proj_bias", (3 * c,)),
                          ("cross_attn.out_proj.weight", (c, c)), ("cross_attn.out_proj.bias", (c,)),
                          ("ln_3.weight", (c,)), ("ln_3.bias", (c,))]
            for n, shp in names:
                if n.endswith(".bias") and not spec["bias"] and "cross_attn" not in n:
                    continue
                out[lp + n] = (shp, f32)
        out[dp + "ln_f.weight"] = ((c,), f32)
        if spec["bias"]:
            out[dp + "ln_f.bias"] = ((c,), f32)
        out["decoder.lm_head.weight"] = ((v, c), f32)
    else:
        dp = "decoder.backbone.transformer."
        out[dp + "wte.weight"] = ((v, c), f32)
        out[dp + "wpe.weight"] = ((1024, c), f32)
        for i in range(L):
            lp = f"{dp}h.{i}."
            names = [("ln_1.weight", (c,)), ("ln_1.bias", (c,)),
                     ("attn.c_attn.weight", (c, 3 * c)), ("attn.c_attn.bias", (3 * c,)),
                     ("attn.c_proj.weight", (c, c)), ("attn.c_proj.bias", (c,)),
                     ("ln_2.weight", (c,)), ("ln_2.bias", (c,))]
            if spec["is_cross_attn"]:
                names += [("crossattention.c_attn.weight", (c, 2 * c)), ("crossattention.c_attn.bias", (2 * c,)),
                          ("crossattention.q_attn.weight", (c, c)), ("crossattention.q_attn.bias", (c,)),
                          ("crossattention.c_proj.weight", (c, c)), ("crossattention.c_proj.bias", (c,)),
                          ("ln_cross_attn.weight", (c,)), ("ln_cross_attn.bias", (c,))]
            names += [("mlp.c_fc.weight", (c, 4 * c)), ("mlp.c_fc.bias", (4 * c,)),
                      ("mlp.c_proj.weight", (4 * c, c)), ("mlp.c_proj.bias", (c,))]
            for n, shp in names:
                out[lp + n] = (shp, f32)
        out[dp + "ln_f.weight"] = ((c,), f32)
        out[dp + "ln_f.bias"] = ((c,), f32)
        out["decoder.backbone.lm_head.weight"] = ((v, c), f32)
    return out


TIED_KEYS = (("decoder.lm_head.weight", "decoder.transformer.wte.weight"),
             ("decoder.backbone.lm_head.weight", "decoder.backbone.transformer.wte.weight"))


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(spec: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic, non-degenerate weights for every schema key (CPU fp32).

    Scales follow the reference initialisers loosely (N(0, 0.02) linears/embeddings,
    ``models/decoder.py:206-212``; N(0,1) EmbeddingBag tables; unit-column LSH projections,
    ``models/layers.py:119-123``) but LayerNorm gains/biases and linear biases are perturbed away
    from 1/0 so that a kernel that drops one of them fails parity.
    """
    sd: Dict[str, torch.Tensor] = {}
    for key, (shape, dtype) in state_schema(spec).items():
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("peer_proj_wt"):
            t = torch.zeros(shape) if len(shape) == 1 else torch.randn(shape, generator=g) / (shape[0] ** 0.5)
        elif ".peer." in key:                    # O(1) scores / activations so that the top-k choice and the softmax are not degenerate
            t = torch.randn(shape, generator=g) * (0.5 if ".emb_" in key else 1.0 / (shape[-1] ** 0.5))
        elif leaf == "projection_mat":
            t = torch.nn.functional.normalize(torch.randn(shape, generator=g), p=2.0, dim=0)
        elif leaf == "grid":
            nb = shape[0]
            t = torch.linspace(-1, 1, nb + 1)[:-1] + 0.5 * (2.0 / nb)
        elif leaf == "pos_offset":
            nb1 = spec["lsh_num_bins"][int(key.split(".emb.")[1].split(".")[0])] + 1
            t = (nb1 * torch.arange(0, shape[0], dtype=torch.long)).reshape(shape)
        elif ".lsh_emb." in key and key.endswith("emb.weight"):
            t = torch.randn(shape, generator=g)
        elif len(shape) == 1 and (".ln" in key or "ln_" in key) and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf in ("class_token", "pos_embedding"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif ".proj.models." in key or key == "encoder.1.weight":
            t = torch.randn(shape, generator=g) / (shape[-1] ** 0.5)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[key] = t.to(dtype).contiguous()
    for a, b in TIED_KEYS:
        if a in sd and b in sd:
            sd[a] = sd[b]
    return sd
