"""KV-cached caption decode: static buffers + one captured CUDA graph per step shape.

What the reference does per new token (models/vision_encoder_decoder.py:144-180): a FULL decoder forward over the whole
prefix, logits for every position, a python n-gram processor with a host sync, top-k / softmax / multinomial / cat.
What this engine does per new token: ~8 fused kernels per layer that touch every decoder weight exactly once
(weight-streaming, HBM-bound), an in-place KV-cache append, an on-device sampler that also advances the position
counter -- the whole step is one CUDA-graph replay with no host round trip.

The result is the same sequence of token ids: text rows never see the soft-prompt rows, so the last-row
logits of the cache-less forward equal the incremental ones (tests/test_gpu_generate.py checks bit-exact greedy ids
against the oracle and the reference-made golden ids).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import call
from .model_spec import layer_has_cross_attn
from .ops import ptr, stream


class DecodeEngine:
    def __init__(self, model, batch: int):
        spec = model.spec
        if spec["decoder"] != "transformer":
            raise NotImplementedError("KV-cached decode is built for TransformerDecoder only (HF GPT-2 layout: next)")
        self.model, self.spec, self.B = model, spec, batch
        self.cd = model.compute_dtype
        dev = next(model.parameters()).device
        self.dev = dev
        C, L, V = spec["n_embd"], spec["n_layer"], spec["vocab_size"]
        self.n_prompt = spec["n_cls"] if spec["use_soft_prompting"] else 0
        self.Tmax = spec["block_size"] - self.n_prompt
        self.ids = torch.zeros((batch, self.Tmax + 1), device=dev, dtype=torch.int64)
        self.pos = torch.zeros(1, device=dev, dtype=torch.int32)
        self.ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        self.kcache = torch.zeros((L, batch, self.Tmax, C), device=dev, dtype=self.cd)
        self.vcache = torch.zeros((L, batch, self.Tmax, C), device=dev, dtype=self.cd)
        self.cross_layers = [d for d in range(L) if spec["use_cross_attn"] and layer_has_cross_attn(spec, d)
                             and (d % 2 == 0 or not spec["skip_alternate_cross_attn"])]
        S = spec["n_cls"]
        self.S = S
        self.xkv = torch.zeros((max(1, len(self.cross_layers)), batch * S, 2 * C), device=dev, dtype=self.cd)
        f32 = dict(device=dev, dtype=torch.float32)
        self.x = torch.zeros((batch, C), **f32)
        self.q = torch.zeros((batch, C), **f32)
        self.y = torch.zeros((batch, C), **f32)
        self.h = torch.zeros((batch, int(spec["ff_mult"] * C)), **f32)
        self.logits = torch.zeros((batch, V), **f32)
        self.ngrams = torch.tensor(list(spec["no_repeat_n_grams"]) or [0], device=dev, dtype=torch.int32)
        self.n_ngrams = len(spec["no_repeat_n_grams"])
        self.seed_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.graphs = {}
        self.launches_per_step = None

    # one decode step for the token at position *pos (device).  sample=False: teacher-forced prompt token.
    def _step(self, sample: bool, temperature: float, top_k: Optional[int], seed: int):
        m, spec, B = self.model, self.spec, self.B
        W = m.weights()
        C, H = spec["n_embd"], spec["n_head"]
        hs = C // H
        wd = ops.F32 if self.cd == torch.float32 else ops.BF16
        st = stream()
        dp = "decoder.transformer."
        pos = ptr(self.pos)
        cbs = self.Tmax * C
        call("i2t_dec_embed", ptr(self.ids), ptr(W[dp + "wte.weight"]), ptr(W[dp + "wpe.weight"]), ptr(self.x), pos, B, C,
             self.ids.shape[1], self.n_prompt, st)

        def lin(x, ln, wkey, bkey, out, ldo, N, K, act=ops.ACT_NONE, residual=None, qkv=None, rows=None):
            w = W.c(wkey)
            b = W.get(bkey) if bkey else None
            if rows is not None:
                w = w[rows]
                b = b[rows] if b is not None else None
            g = W[ln + ".weight"] if ln else None
            be = W.get(ln + ".bias") if ln else None
            kc, vc = (ptr(qkv[0]), ptr(qkv[1])) if qkv else (None, None)
            call("i2t_dec_linear", ptr(x), ptr(g), ptr(be), 1e-5, ptr(w), ptr(b), ptr(residual), ptr(out), ldo, B, N, K, act,
                 wd, 1 if qkv else 0, kc, vc, cbs, C, wd, pos, st)

        xi = 0
        for d in range(spec["n_layer"]):
            lp = f"{dp}h.{d}."
            lin(self.x, lp + "ln_1", lp + "attn.c_attn.weight", lp + "attn.c_attn.bias", self.q, C, 3 * C, C,
                qkv=(self.kcache[d], self.vcache[d]))
            call("i2t_dec_attn", ptr(self.q), C, ptr(self.kcache[d]), ptr(self.vcache[d]), cbs, C, ptr(self.y), C, pos, 1, B, H,
                 hs, wd, st)
            lin(self.y, None, lp + "attn.c_proj.weight", lp + "attn.c_proj.bias", self.x, C, C, C, residual=self.x)
            if d in self.cross_layers:
                kw, kb = lp + "cross_attn.in_proj_weight", lp + "cross_attn.in_proj_bias"
                lin(self.x, lp + "ln_3", kw, kb, self.q, C, C, C, rows=slice(0, C))
                kv = self.xkv[xi]
                es = kv.element_size()
                call("i2t_dec_attn", ptr(self.q), C, kv.data_ptr(), kv.data_ptr() + C * es, self.S * 2 * C, 2 * C, ptr(self.y),
                     C, None, self.S, B, H, hs, wd, st)
                lin(self.y, None, lp + "cross_attn.out_proj.weight", lp + "cross_attn.out_proj.bias", self.x, C, C, C,
                    residual=self.x)
                xi += 1
            F = self.h.shape[1]
            lin(self.x, lp + "ln_2", lp + "mlp.c_fc.weight", lp + "mlp.c_fc.bias", self.h, F, F, C, act=ops.ACT_GELU_TANH)
            lin(self.h, None, lp + "mlp.c_proj.weight", lp + "mlp.c_proj.bias", self.x, C, C, F, residual=self.x)
        if sample:
            V = spec["vocab_size"]
            lin(self.x, dp + "ln_f", "decoder.lm_head.weight", None, self.logits, V, V, C)
            call("i2t_sample", ptr(self.logits), V, B, V, ptr(self.ids), self.ids.shape[1], pos, 1, 0, temperature,
                 int(top_k) if top_k is not None else 0, ptr(self.ngrams), self.n_ngrams, 0, ptr(self.seed_dev), None, ptr(self.ticket), 1, st)
        else:
            call("i2t_dec_advance", pos, st)

    def _prefill_cross(self, enc: torch.Tensor):
        """K/V projections of the (fixed) encoder output, once per generate call: the k,v rows of
        cross_attn.in_proj applied to the RAW encoder output (reference models/layers.py:600-605)."""
        W = self.model.weights()
        C = self.spec["n_embd"]
        e = enc.reshape(self.B * self.S, C)
        e = e if e.dtype == self.cd else e.to(self.cd)
        for xi, d in enumerate(self.cross_layers):
            lp = f"decoder.transformer.h.{d}."
            w = W.c(lp + "cross_attn.in_proj_weight")[C:]
            b = W[lp + "cross_attn.in_proj_bias"][C:]
            ops.gemm(e, w, bias=b, out=self.xkv[xi])

    @torch.no_grad()
    def generate(self, images, prompt_ids, max_new_tokens: int, temperature: float, top_k: Optional[int], seed: int):
        from . import functional as Fn
        from ._lib import launch_count
        m, B = self.model, self.B
        P = prompt_ids.shape[1]
        assert prompt_ids.shape[0] == B and P + max_new_tokens <= self.Tmax + 1
        enc = Fn.encoder_forward(m.weights(), self.spec, images, self.cd, train_trunk=False)
        self._prefill_cross(enc)
        self.ids[:, :P].copy_(prompt_ids)
        self.pos.zero_()
        self.ticket.zero_()
        for _ in range(P - 1):
            self._step(False, temperature, top_k, seed)
        self.seed_dev.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        key = (float(temperature), top_k)
        g = self.graphs.get(key)
        if g is None:
            # first use: run one step eagerly (sets kernel attributes, loads modules), then capture the next one
            n0 = launch_count()
            self._step(True, temperature, top_k, seed)
            self.launches_per_step = launch_count() - n0
            done = 1
            if max_new_tokens > 1:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step(True, temperature, top_k, seed)
                # capture does not execute: the graph now holds exactly one step
                if len(self.graphs) > 16:
                    self.graphs.clear()
                self.graphs[key] = g
        else:
            done = 0
        self.replays_last = max_new_tokens - done
        for _ in range(max_new_tokens - done):
            g.replay()
        return self.ids[:, :P + max_new_tokens].clone()
