"""Beam-search caption generation over the B200 model: same constructor, call signature and results as the reference's
``models.generation_utils.BeamSearchTokenGenerator`` (models/generation_utils.py:10-148).

Per step (TransformerDecoder): ONE KV-cached decode step over all beam_width x batch hypothesis rows (DecodeEngine, a CUDA-graph
replay; the reference re-runs the whole decoder over the prefix, models/generation_utils.py:65); after the consolidation the cached
K / V rows and the token history are reordered by the surviving beams ON THE DEVICE.  HF-layout decoders keep the cache-less forward.
The device sampler kernel turns the last-position logits into the post-ban / post-top-k distribution (no-repeat-n-gram ban and
top-k threshold of the reference's `decode_next`), each hypothesis proposes `beam_expansion_factor` continuations (arg-top
when temperature <= 0, multinomial otherwise), and the beam_width best (or sampled, `consolidation_temperature` > 0) of
the beam_width * expansion candidates per image survive.  Hypotheses that already emitted EOS keep emitting EOS at zero
cost unless a continuation still beats the length boost -- the reference's rule.  Book-keeping tensors are tiny
(beam_width x batch x expansion) and stay on the device.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import ops
from ._lib import call


class BeamSearchTokenGenerator:
    def __init__(self, model, beam_width: int = 3, temperature: float = 1.0, top_k: Optional[int] = None, max_new_tokens=64,
                 no_repeat_n_grams: Sequence[int] = (2, 3, 4), beam_expansion_factor: int = 4, eos_token_id: Optional[int] = None,
                 consolidation_temperature: float = 1.0, length_boost: float = 1.0):
        self.model = model
        self.beam_width = beam_width
        self.beam_expansion_factor = beam_expansion_factor
        self.max_new_tokens = max_new_tokens
        self.temperature = temperature
        self.consolidation_temperature = consolidation_temperature
        self.top_k = top_k
        self.eos_token_id = eos_token_id
        self.length_boost = math.log(length_boost)
        self.no_repeat_n_grams = tuple(no_repeat_n_grams)

    # -- next-token log-probabilities of every hypothesis: (rows, V) --------------------------------------------------------
    def _log_probs(self, enc_rows: torch.Tensor, ids_rows: torch.Tensor, engine=None) -> torch.Tensor:
        V = self.model.spec["vocab_size"]
        if engine is not None:                                   # KV-cached: one decode step over the cache
            scores = engine.beam_logits()[:, :V].float().contiguous()
        else:
            out = self.model(images=None, ids=ids_rows, encoder_output=enc_rows)
            scores = out.logits[:, -1, :V].float().contiguous()
        rows, cur = ids_rows.shape
        hist = torch.zeros((rows, cur + 1), device=ids_rows.device, dtype=torch.int64)
        hist[:, :cur] = ids_rows
        ngrams = torch.tensor(list(self.no_repeat_n_grams) or [0], device=ids_rows.device, dtype=torch.int32)
        probs = torch.empty((rows, V), device=ids_rows.device, dtype=torch.float32)
        temp = self.temperature if self.temperature > 0 else 1.0
        call("i2t_sample", ops.ptr(scores), V, rows, V, ops.ptr(hist), hist.shape[1], None, 0, cur, float(temp),
             int(self.top_k) if self.top_k is not None else 0, 0.0, ops.ptr(ngrams), len(self.no_repeat_n_grams), 0, None,
             ops.ptr(probs), None, 0, ops.stream())
        return probs.log()        # -inf where banned / outside the top-k

    @torch.no_grad()
    def __call__(self, inputs: torch.Tensor, decoded_ids: torch.Tensor):
        self.model.eval()
        bw, ex = self.beam_width, self.beam_expansion_factor
        bs = inputs.size(0)
        enc = self.model.encoder(inputs)                                   # (bs, n_cls, C)
        enc_rows = enc.unsqueeze(0).expand(bw, -1, -1, -1).reshape(bw * bs, enc.size(1), enc.size(2)).contiguous()
        provided = decoded_ids.size(-1) - 1
        beams = decoded_ids.unsqueeze(0).expand(bw, -1, -1).contiguous()    # (bw, bs, L) beam-major like the reference
        engine = None
        spec = self.model.spec
        blk = spec["block_size"] - (spec["n_cls"] if spec["use_soft_prompting"] else 0)
        if spec["decoder"] == "transformer" and self.max_new_tokens + provided <= blk and not getattr(self, "cacheless", False):
            from .decode_engine import DecodeEngine
            key = ("beam", bw * bs, self.model.compute_dtype)
            engine = self.model._decode_engines.get(key)
            if engine is None:
                engine = self.model._decode_engines[key] = DecodeEngine(self.model, bw * bs, mode="kernels")
            engine.beam_begin(enc_rows, beams.reshape(bw * bs, -1))
        total = torch.zeros((bw, bs), device=enc.device)
        while beams.size(-1) < self.max_new_tokens + provided:
            if self.eos_token_id is not None and bool(((beams == self.eos_token_id).sum(dim=-1) > 0).all()):
                break
            L = beams.size(-1)
            rows = beams.reshape(bw * bs, L)
            logp = self._log_probs(enc_rows, rows, engine)
            if self.temperature <= 0:
                nxt = logp.topk(k=ex, dim=-1, sorted=False).indices
            else:
                nxt = torch.multinomial(logp.exp(), num_samples=ex)
            step = torch.gather(logp, -1, nxt)
            if self.eos_token_id is not None:
                done = (rows[:, -1] == self.eos_token_id).unsqueeze(-1)
                stay = torch.logical_and(done, step + self.length_boost < 0)
                nxt = torch.where(stay, torch.full_like(nxt, self.eos_token_id), nxt)
                step = torch.where(stay, torch.zeros_like(step), step + self.length_boost)
            nxt = nxt.reshape(bw, bs, ex)
            step = step.reshape(bw, bs, ex)
            # candidates of one image: (beam, expansion) pairs flattened beam-major
            cand = (total.unsqueeze(2) + step).permute(1, 0, 2).reshape(bs, bw * ex)
            if self.consolidation_temperature <= 0:
                pick = cand.topk(k=bw, dim=-1, sorted=True).indices
            else:
                pick = torch.multinomial((cand / self.consolidation_temperature).softmax(dim=-1), num_samples=bw)
            src_beam, src_exp = pick // ex, pick % ex                          # (bs, bw)
            b_idx = torch.arange(bs, device=enc.device).unsqueeze(1).expand(-1, bw)
            kept = beams[src_beam, b_idx]                                      # (bs, bw, L)
            tok = nxt[src_beam, b_idx, src_exp]                                # (bs, bw)
            beams = torch.cat((kept, tok.unsqueeze(-1)), dim=-1).permute(1, 0, 2).contiguous()
            if engine is not None:             # new row (j, b) continues old row (src_beam[b, j], b): reorder the cache on the device
                src_rows = (src_beam * bs + b_idx).permute(1, 0).reshape(-1)
                engine.beam_advance(src_rows, tok.permute(1, 0).reshape(-1), L)
            total = (total[src_beam, b_idx] + step[src_beam, b_idx, src_exp]).permute(1, 0).contiguous()
        return beams.permute(1, 0, 2), total.permute(1, 0)
