"""Drop-in for the reference's ``models.vision_encoder_decoder.VisionEncoderDecoder`` on B200.

Same constructor, ``forward`` / ``generate`` signatures, return type and ``state_dict`` layout as the reference
(models/vision_encoder_decoder.py:17-182); every arithmetic step runs in libi2t (hand-written sm_100a CUDA).  There is
no eager / CPU fallback: tensors must live on a CUDA device.

Behavioural notes that parity depends on (all pinned by tests/golden):
  * D9 -- the reference turns the caller's ``attn_msk`` into an all-zero float mask
    (vision_encoder_decoder.py:101-102,117-118), so ``attn_msk`` never changes an output.  It is accepted and ignored.
  * Q1 -- with soft prompting the text rows never see the prompt rows; the prompt only shifts positions.
  * greedy decoding is ``top_k=1``; the no-repeat-n-gram ban runs every step.
"""
from __future__ import annotations

from collections import namedtuple
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .config_schema import VisionEncoderDecoderConfig
from .model_spec import TIED_KEYS, layer_has_cross_attn, spec_from_config, state_schema
from .modules import ParamTree, build_param_tree

VisionEncoderDecoderModelOutput = namedtuple("VisionEncoderDecoderModelOutput",
                                             ["encoder_output", "logits", "hidden_state"])


class _Side(ParamTree):
    """`.encoder` / `.decoder` attribute of the model: parameter container + the properties the reference exposes."""

    def __init__(self, kind: str, spec: dict):
        super().__init__()
        self._kind, self._spec = kind, spec

    # Encoder API (reference models/encoder.py:47-53,121-127)
    @property
    def num_outputs(self):
        return self._spec["n_cls"]

    @property
    def output_embed_dim(self):
        return self._spec["n_embd_out_vit"]

    # Decoder API (reference models/decoder.py:150-157,262-281)
    @property
    def block_size(self):
        return self._spec["block_size"]

    @property
    def n_embd(self):
        return self._spec["n_embd"]

    # Callable like the reference's sub-modules (BeamSearchTokenGenerator calls `model.encoder(images)`,
    # models/generation_utils.py:37; the decoder API is models/decoder.py:214-256,150-157).
    def _bind(self, owner):
        import weakref
        self.__dict__["_owner"] = weakref.ref(owner)      # not a registered sub-module: no cycle in state_dict()

    def forward(self, *args, **kwargs):
        from . import functional as Fn
        m = self.__dict__["_owner"]()
        W = m.weights()
        if self._kind == "encoder":
            images = args[0] if args else kwargs["images"]
            return Fn.encoder_forward(W, self._spec, images, m.compute_dtype,
                                      train_trunk=m.training and self._spec["refine_base_model"])
        idx = kwargs.get("idx", args[0] if args else None)
        embeds = kwargs.get("inputs_embeds", args[1] if len(args) > 1 else None)
        cross = kwargs.get("cross_attn_embeds", args[2] if len(args) > 2 else None)
        if embeds is None and idx is None:
            raise ValueError("Decoder.forward needs idx or inputs_embeds")
        # stand-alone decoder call (reference models/decoder.py:214-256): rows are wte[idx] or the given inputs_embeds, + wpe;
        # no soft-prompt semantics, cross attention iff embeddings are given; the blocks' own causal mask applies (a caller's
        # additive attn_msk is not: the only caller in the reference passes the closed forms of Appendix C, which
        # VisionEncoderDecoder.forward composes itself)
        spec = dict(self._spec, use_soft_prompting=False, use_cross_attn=cross is not None)
        lead = embeds if embeds is not None else idx
        enc = cross if cross is not None else torch.zeros((lead.shape[0], 1, spec["n_embd"]), device=lead.device)
        logits, hidden = Fn.decoder_forward(W, spec, idx, enc, m.compute_dtype, training=m.training, drop=m._drop_ctx(),
                                            inputs_embeds=embeds)
        return logits[..., :spec["vocab_size"]], hidden

    def get_inputs_embeds(self, idx):
        m = self.__dict__["_owner"]()
        key = "decoder.transformer.wte.weight" if self._spec["decoder"] == "transformer" else "decoder.backbone.transformer.wte.weight"
        return m.weights()[key][idx]

    def tie_weights(self):
        """lm_head.weight and wte.weight are ONE storage by construction (both keys appear in the state dict)."""
        return None


class VisionEncoderDecoder(nn.Module):
    def __init__(self, config: VisionEncoderDecoderConfig, encoder=None, decoder=None, spec_overrides: Optional[dict] = None,
                 device="cuda", compute_dtype: torch.dtype = torch.float32, seed: Optional[int] = None):
        super().__init__()
        if decoder is not None:
            raise NotImplementedError("constructor injection of a foreign DECODER is not supported: the decoder is the B200 "
                                      "hot path (its kernels own the weight layout and the KV cache); inject an encoder instead")
        if not (config.use_cross_attn or config.use_soft_prompting):
            raise ValueError("Misconfigured!!! Need to either use cross attn or soft prompting or both")
        self.config = config
        self.spec = spec_from_config(config, **(spec_overrides or {}))
        if encoder is not None:
            # reference models/vision_encoder_decoder.py:19-37: a caller-supplied Encoder (any nn.Module with `num_outputs`,
            # `output_embed_dim` and forward(images) -> (B, num_outputs, output_embed_dim)) replaces Encoder.from_config; it
            # runs as the torch module it is and feeds `encoder_output` to the B200 decoder path
            self.spec = dict(self.spec, n_cls=int(encoder.num_outputs), n_embd_out_vit=int(encoder.output_embed_dim),
                             injected_encoder=True)
        spec = self.spec
        if spec["decoder"] == "transformer" and spec["use_soft_prompting"] and not spec["is_causal"]:
            raise NotImplementedError("non-causal decoder blocks with a soft prompt are a 'next' row (SURVEY.md 8f-1)")
        if spec["decoder"] == "transformer" and spec["use_cross_attn"] != spec["is_cross_attn"]:
            raise ValueError("use_cross_attn must match decoder transformer_config.is_cross_attn")
        self.compute_dtype = compute_dtype
        gen = None
        if seed is not None:
            gen = torch.Generator(device="cpu").manual_seed(seed)
        schema = state_schema(spec)
        if encoder is None:                     # (the encoder registers first: state_dict / named_parameters order of the reference)
            self.encoder = _Side("encoder", spec)
            self.encoder._bind(self)
        else:
            schema = type(schema)((k, v) for k, v in schema.items() if not k.startswith("encoder"))
            if spec["n_embd_out_vit"] != spec["n_embd"]:            # the reference bridges with a Linear (:33-37)
                encoder = nn.Sequential(encoder, nn.Linear(spec["n_embd_out_vit"], spec["n_embd"], bias=False))
            self.encoder = encoder.to(torch.device(device))
        self.decoder = _Side("decoder", spec)
        self.decoder._bind(self)
        build_param_tree(self, schema, spec, TIED_KEYS, torch.device(device), gen)
        self.space_for_prompt = spec["n_cls"] if config.use_soft_prompting else 0
        self.use_cross_attn = config.use_cross_attn
        self.use_soft_prompting = config.use_soft_prompting
        self.processor = tuple(spec["no_repeat_n_grams"])   # consumed on the device by the sampler kernel
        self._ptr_tables = None
        self._decode_engines = {}
        self._rng = None               # device int64[2] {seed, step offset} of the training-mode dropout masks
        self._drop_seed = 0x1234ABCD if seed is None else int(seed)
        if config.chkpt_path is not None:
            self.load_partial_checkpoint(config.chkpt_path)

    # ------------------------------------------------------------------ parameters ----------------------------
    def __setstate__(self, state):
        super().__setstate__(state)            # copy.deepcopy / pickle: re-point the sub-objects at THIS model
        self.__dict__["_tensor_cache"] = None
        if isinstance(self.encoder, _Side):
            self.encoder._bind(self)
        self.decoder._bind(self)

    def encode(self, images: torch.Tensor) -> torch.Tensor:
        """images -> (B, n_cls, n_embd) fp32 encoder output: the B200 ViT path, or the injected torch encoder."""
        from . import functional as Fn
        if self.spec.get("injected_encoder"):
            return self.encoder(images).float()
        return Fn.encoder_forward(self.weights(), self.spec, images, self.compute_dtype,
                                  train_trunk=self.training and self.spec["refine_base_model"])

    def load_partial_checkpoint(self, path: str, map_location=None):
        """reference models/utils.py:31-36: state_dict().update(torch.load(path)); load_state_dict."""
        full = self.state_dict()
        full.update(torch.load(path, map_location=map_location))
        self.load_state_dict(full)
        return self

    def _tensors(self) -> Dict[str, torch.Tensor]:
        """name -> Parameter / buffer.  Memoised: walking the 432-tensor module tree costs ~0.3 ms of host time, which every
        forward() and generate() call used to pay (twice) while the GPU sat idle.  The Parameter objects are stable under
        optimiser steps, load_state_dict (in-place copies) and .to(); the cache is dropped wherever they could be replaced."""
        c = self.__dict__.get("_tensor_cache")
        if c is None:
            c = dict(self.named_parameters(remove_duplicate=False))
            c.update(dict(self.named_buffers()))
            self.__dict__["_tensor_cache"] = c
        return c

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_tensor_cache"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__["_tensor_cache"] = None
        return super().load_state_dict(*args, **kwargs)

    def register_parameter(self, name, param):
        self.__dict__["_tensor_cache"] = None
        return super().register_parameter(name, param)

    def _lsh_tables(self, pre: str):
        """Device pointer tables for the LSH tail kernel (slot-major, then resolution)."""
        t = self._tensors()
        spec = self.spec
        keys = [(s, r) for s in range(spec["n_cls"]) for r in range(len(spec["lsh_num_bins"]))]
        ptrs = {k: [t[f"{pre}lsh_emb.{s}.emb.{r}.{leaf}"].data_ptr() for s, r in keys]
                for k, leaf in (("proj", "projection_mat"), ("grid", "grid"), ("emb", "emb.weight"))}
        sig = tuple(ptrs["emb"]) + tuple(ptrs["proj"])
        if self._ptr_tables is None or self._ptr_tables[0] != sig:
            dev = t[f"{pre}lsh_emb.0.emb.0.emb.weight"].device
            tabs = {k: torch.tensor(v, dtype=torch.int64, device=dev) for k, v in ptrs.items()}
            tabs["nb"] = torch.tensor(list(spec["lsh_num_bins"]), dtype=torch.int32, device=dev)
            self._ptr_tables = (sig, tabs)
        return self._ptr_tables[1]

    # ------------------------------------------------------------------ dropout ------------------------------
    def set_dropout_seed(self, seed: int):
        """Seed of the counter-based dropout masks (csrc/rng.cuh); the step offset restarts at 0."""
        self._drop_seed = int(seed)
        self._rng = None

    def _drop_ctx(self):
        """Mask context of ONE training-mode forward pass, or None (eval mode / every p == 0).  The shared step offset is
        bumped on the stream and snapshotted, so the autograd nodes of this pass regenerate exactly its masks however many
        other passes run before their backward (and a CUDA-graph replay of forward + backward draws fresh masks)."""
        from . import ops
        if not self.training or (self.spec.get("dropout", 0.0) <= 0.0 and self.spec.get("attn_dropout", 0.0) <= 0.0):
            return None
        dev = next(self.parameters()).device
        if self._rng is None or self._rng.device != dev:
            self._rng = torch.tensor([self._drop_seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        ops.rng_advance(self._rng)
        return ops.DropCtx(self._rng.clone())

    def weights(self):
        """key -> tensor accessor; `.c(key)` returns the tensor in the compute dtype.  bf16 copies live on the Parameter
        object (`_i2t_shadow`) and are refreshed IN PLACE when the fp32 master's version counter moved (load_state_dict,
        torch ops) -- the fused optimiser / EMA kernels refresh them themselves (optimizer.py) -- so their addresses stay
        valid for captured CUDA graphs and decode tables."""
        return _Weights(self)

    @torch.no_grad()
    def sync_compute_weights(self):
        """Refresh every stale bf16 copy now (a CUDA-graph replay runs no Python, so nothing would do it lazily)."""
        if self.compute_dtype == torch.float32:
            return
        stale = [w for w in self._tensors().values()
                 if getattr(w, "_i2t_shadow", None) is not None and w._i2t_shadow_version != w._version]
        seen = set()
        for w in stale:
            if id(w) in seen:
                continue
            seen.add(id(w))
            w._i2t_shadow.copy_(w.detach())
            w._i2t_shadow_version = w._version

    def weight_generation(self):
        """Changes whenever a decoder weight changed (version counters; the fused optimiser bumps them)."""
        c = self.__dict__.get("_wgen_tensors")
        t = self._tensors()
        if c is None or c[0] is not t:
            c = (t, [w for k, w in t.items() if k.startswith("decoder.") and w.dim() == 2])
            self.__dict__["_wgen_tensors"] = c
        return tuple(w._version for w in c[1]) + tuple(w.data_ptr() for w in c[1][:1])

    # ------------------------------------------------------------------ forward -------------------------------
    def forward(self, images: Optional[torch.Tensor], ids: torch.Tensor, attn_msk: Optional[torch.Tensor] = None,
                encoder_output: Optional[torch.Tensor] = None, _padded_logits: bool = False) -> VisionEncoderDecoderModelOutput:
        from . import functional as Fn
        W = self.weights()
        if encoder_output is None:
            encoder_output = self.encode(images)
        # attn_msk: accepted and ignored -- it has no effect in the reference either (D9)
        logits, hidden = Fn.decoder_forward(W, self.spec, ids, encoder_output, self.compute_dtype, training=self.training,
                                            drop=self._drop_ctx())
        if not _padded_logits:
            logits = logits.contiguous()      # the reference returns contiguous logits (vision_encoder_decoder.py:132);
            # the trainer wrapper asks for the row-padded view instead so the LM-head backward runs in place
        return VisionEncoderDecoderModelOutput(encoder_output=encoder_output, logits=logits, hidden_state=hidden)

    @torch.no_grad()
    def generate(self, images, prompt_ids, max_new_tokens=128, temperature=1.0, top_k=None, nucleus_p=None,
                 seed: Optional[int] = None) -> torch.LongTensor:
        """reference models/vision_encoder_decoder.py:136-182, with a KV cache, an on-device sampler and one launch (<= 8
        sequences, bf16) or one CUDA graph per step shape.  Returns (B, prompt + max_new_tokens) int64 including the prompt.
        Determinism: fp32 greedy ids are bit-exact and run-to-run identical (fixed summation order); bf16 with <= 8 sequences
        (the dataflow megakernel) is run-to-run identical too; bf16 with MORE than 8 sequences adds split-K partial tiles
        with fp32 atomics, so greedy ids may differ between runs at near-ties -- `image2text_b200._lib.lib().
        i2t_set_gemm_split_k(0)` (or I2T_GEMM_SPLITK=0) selects the fixed-order projections at ~20 % lower tok/s."""
        from .decode_engine import DecodeEngine
        blk = self.spec["block_size"] - self.space_for_prompt
        assert max_new_tokens <= blk - prompt_ids.size(-1)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # torch's RNG seeds the device Philox stream
        if nucleus_p is not None and float(nucleus_p) <= 0.0:
            # the reference keeps the sorted tokens whose cumulative probability is <= max(p, p_max): p <= 0 leaves only the
            # most probable one (vision_encoder_decoder.py:160-172) -- the greedy pick after the ban / top-k filter
            top_k, nucleus_p = 1, None
        if nucleus_p is not None and float(nucleus_p) >= 1.0:
            nucleus_p = None                   # keeps everything: plain sampling
        if self.spec["decoder"] != "transformer":
            if os.environ.get("I2T_HF_DECODE", "cached") == "cacheless":
                return self._generate_cacheless(images, prompt_ids, max_new_tokens, float(temperature), top_k, nucleus_p, seed)
            from .decode_engine import HFDecodeEngine
            need = self.space_for_prompt + prompt_ids.shape[1] + max_new_tokens
            key = ("hf", prompt_ids.shape[0], self.compute_dtype)
            eng = self._decode_engines.get(key)
            if eng is None or eng.cap < need:
                eng = HFDecodeEngine(self, prompt_ids.shape[0], capacity=max(need, 128))
                self._decode_engines[key] = eng
            return eng.generate(images, prompt_ids, max_new_tokens, float(temperature), top_k, seed, nucleus_p=nucleus_p)
        B = prompt_ids.shape[0]
        key = (B, self.compute_dtype, nucleus_p is not None)
        eng = self._decode_engines.get(key)
        if eng is None:
            # the top-p filter lives in the stand-alone sampler kernel: that request takes the per-stage kernels
            eng = DecodeEngine(self, B, mode="kernels" if nucleus_p is not None else None)
            self._decode_engines[key] = eng
        return eng.generate(images, prompt_ids, max_new_tokens, float(temperature), top_k, seed, nucleus_p=nucleus_p)

    @torch.no_grad()
    def _generate_cacheless(self, images, prompt_ids, max_new_tokens, temperature, top_k, nucleus_p, seed):
        """HF-layout decoders (GPT2HuggingfaceDecoder, models/decoder.py:335-361): the reference's own algorithm -- one full
        decoder forward over the prefix per new token (vision_encoder_decoder.py:144-150) -- on the CUDA kernels, with the
        device sampler (n-gram ban, top-k, top-p, Philox draw) instead of the host loop of :152-180.  The encoder runs once.
        A KV-cached decode for the Conv1D weight layout is the next row; this path exists so that `generate` works for every
        model the forward supports."""
        from . import functional as Fn
        from . import ops
        from ._lib import call
        spec = self.spec
        W = self.weights()
        enc = self.encode(images)
        B, P = prompt_ids.shape
        V = spec["vocab_size"]
        blk = spec["block_size"] - self.space_for_prompt
        dev = prompt_ids.device
        ids = torch.zeros((B, P + max_new_tokens), device=dev, dtype=torch.int64)
        ids[:, :P] = prompt_ids
        ngrams = torch.tensor(list(spec["no_repeat_n_grams"]) or [0], device=dev, dtype=torch.int32)
        n_ngrams = len(spec["no_repeat_n_grams"])
        row = torch.empty((B, V), device=dev, dtype=torch.float32)
        for t in range(max_new_tokens):
            cur = P + t
            cond = ids[:, max(0, cur - blk):cur].contiguous()
            logits, _ = Fn.decoder_forward(W, spec, cond, enc, self.compute_dtype, training=False)
            row.copy_(logits[:, -1, :V])
            call("i2t_sample", ops.ptr(row), V, B, V, ops.ptr(ids), ids.shape[1], None, 0, cur, temperature,
                 int(top_k) if top_k is not None else 0, float(nucleus_p or 0.0), ops.ptr(ngrams), n_ngrams,
                 int(seed) & 0x7FFFFFFFFFFFFFFF, None, None, None, 1, ops.stream())
        return ids


class _Weights:
    def __init__(self, model: VisionEncoderDecoder):
        self.m = model
        self.t = model._tensors()

    def __contains__(self, key):
        return key in self.t

    def __getitem__(self, key) -> torch.Tensor:
        return self.t[key]

    def get(self, key):
        return self.t.get(key)

    def c(self, key, rows: Optional[slice] = None) -> torch.Tensor:
        w = self.t[key]
        cd = self.m.compute_dtype
        if cd != torch.float32:
            sh = getattr(w, "_i2t_shadow", None)
            if sh is None or sh.dtype != cd or sh.device != w.device or sh.shape != w.shape:
                sh = w.detach().to(cd).contiguous()
                w._i2t_shadow, w._i2t_shadow_version = sh, w._version
            elif w._i2t_shadow_version != w._version:
                sh.copy_(w.detach())                   # in place: captured graphs / decode tables keep the address
                w._i2t_shadow_version = w._version
            w = sh
        return w if rows is None else w[rows]
