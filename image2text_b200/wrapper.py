"""Mirror of the reference's ``training.wrapper.ModelTrainerWrapper`` (training/wrapper.py:13-214) over the B200 model.

Same constructor arguments, same ``train_step`` / ``val_step`` / ``forward`` / ``forward_m`` / ``copy_momentum_params``
surface, same loss definitions:
  * token corruption (MLM) and BOS shift are integer bookkeeping on the id tensors (training/wrapper.py:154-196);
  * the LM loss -- weighted CE, or the momentum-distillation soft-CE -- and its gradient w.r.t. the logits come from ONE
    fused pass over the vocabulary (``i2t_lm_loss``), never materialising the (B,T,V+1) one-hot of :136-141;
  * the EMA teacher update is one in-place multi-tensor launch (``i2t_ema_multi``) instead of a per-parameter
    allocate-and-rebind loop (:53-60).  It runs every micro-step before backward, like the reference.
The contrastive auxiliary loss (:98-118, off in every reference YAML) is `compute_contrastive_loss` (one similarity GEMM + a
masked cross-entropy kernel).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from .autograd_ops import LmLossFn
from .config_schema import TrainerWrapperConfig, VisionEncoderDecoderConfig
from .optimizer import EmaUpdater
from .vision_encoder_decoder import VisionEncoderDecoder


class ModelTrainerWrapper(nn.Module):
    def __init__(self, model_config: VisionEncoderDecoderConfig, tokenizer, trainer_config: TrainerWrapperConfig,
                 ignore_index: int = -100, device="cuda", compute_dtype: torch.dtype = torch.float32, **model_kw):
        super().__init__()
        self.model = VisionEncoderDecoder(config=model_config, device=device, compute_dtype=compute_dtype, **model_kw)
        self.is_momentum = trainer_config.moco_momentum is not None and trainer_config.moco_alpha is not None
        self.model_m = VisionEncoderDecoder(config=model_config, device=device, compute_dtype=compute_dtype, **model_kw) \
            if self.is_momentum else None
        self.tokenizer = tokenizer
        self.ignore_index = ignore_index
        self.temperature = trainer_config.training_temperature
        self.weight_fn = trainer_config.weight_fn
        if self.weight_fn not in ("constant", "inverse_sqrt_position"):
            raise ValueError(f"unknown weight_fn: {self.weight_fn}")
        self.mask_fraction = trainer_config.mask_fraction
        self.random_mask_fraction = trainer_config.random_mask_fraction
        self.eos_token_weight = trainer_config.eos_token_weight
        self.add_contrastive_loss = trainer_config.add_contrastive_loss
        self.contrastive_temperature = trainer_config.training_contrastive_temperature
        self.momentum = trainer_config.moco_momentum
        self.alpha = trainer_config.moco_alpha
        self._ema = EmaUpdater()
        if self.model_m is not None:          # the teacher runs in train mode too: its own mask stream
            self.model_m.set_dropout_seed(self.model._drop_seed + 1)
        self.copy_momentum_params()

    @torch.no_grad()
    def copy_momentum_params(self):
        if self.is_momentum:
            self.model_m.load_state_dict(self.model.state_dict())      # buffers too (training/wrapper.py:46-51)

    @torch.no_grad()
    def _momentum_update(self):
        if not self.is_momentum:
            return
        ps = [p for _, p in self.model.named_parameters()]
        pm = [p for _, p in self.model_m.named_parameters()]
        self._ema(pm, ps, self.momentum)

    def forward(self, images, input_ids, attn_msk=None) -> Tuple[torch.Tensor, torch.Tensor]:
        out = self.model(images=images, ids=input_ids, attn_msk=attn_msk, _padded_logits=True)
        return out.logits, out.hidden_state

    @torch.no_grad()
    def forward_m(self, images, input_ids, attn_msk=None) -> Tuple[torch.Tensor, torch.Tensor]:
        out = self.model_m(images=images, ids=input_ids, attn_msk=attn_msk, _padded_logits=True)
        return out.logits, out.hidden_state

    def train_step(self, images, labels):
        return self._step(images, labels, True)

    def val_step(self, images, labels):
        return self._step(images, labels, False)

    def train_step_graphed(self, images, labels, loss_scale: float = 1.0, reducer=None, sync: bool = False):
        """train_step + backward of (loss * loss_scale) as ONE CUDA-graph replay (no counterpart in the reference, whose
        loop is `loss, _ = train_step(...); accelerator.backward(loss)`, training/utils.py:76-88).  A micro-step is ~1500
        kernel launches; issued one by one from Python they leave the GPU idle half of the time, so the fixed-shape
        forward + backward is captured once (third call: the first two run eagerly to create the .grad buffers and warm
        up kernel attributes / TMA descriptors) and replayed from static input buffers.  Gradients ACCUMULATE into the
        existing .grad tensors exactly like the eager path (keep `zero_grad(set_to_none=False)`: the graph holds their
        addresses; parameters that received no gradient in the two eager calls are unused and stay untouched).  Returns the detached
        loss (a static tensor that the next replay overwrites).
        Data parallel: pass the `GradientAllReducer` as `reducer` (always) and `sync=True` on the last micro-step of an
        optimiser step; `reducer.finish()` then exchanges the buckets after the replay (NVLink moves them in a few ms).  With
        `GradientAllReducer(graph_events=True)` the captured graph instead carries one external event per gradient bucket and the
        all-reduce of a synchronising replay is queued behind those events, overlapping the rest of the backward -- measured
        slower on B200, because every event node interrupts the replay's launch chain (dp.py)."""
        key = (tuple(images.shape), images.dtype, tuple(labels.shape), float(loss_scale))
        st = getattr(self, "_graph_state", None)
        if st is None or st["key"] != key:
            st = self._graph_state = dict(key=key, calls=0, graph=None)
        st["calls"] += 1
        if st["graph"] is None and (st["calls"] <= 2 or st.get("failed")):
            loss, _ = self._step(images, labels, True)
            (loss * loss_scale).backward()
            return loss.detach()
        if st["graph"] is None:
            st["images"], st["labels"] = images.clone(), labels.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                self._grad_prescale = float(loss_scale)       # folded into dlogits by the loss kernel (no rescale pass)
                import contextlib
                from . import ops
                # ops.grad_sinks: the captured backward adds weight gradients straight into the (static) .grad buffers
                events = reducer is not None and reducer.graph_events
                with (reducer.capturing() if events else contextlib.nullcontext()), \
                        ops.grad_sinks(notify=reducer._hook if events else None):
                    with torch.cuda.graph(g):
                        loss, _ = self._step(st["images"], st["labels"], True)
                        (loss * loss_scale).backward()
                        st["loss"] = loss.detach()
            except Exception:
                st["failed"] = True
                torch.cuda.synchronize()
                raise
            finally:
                self._grad_prescale = 1.0
            st["graph"] = g
        st["images"].copy_(images)
        st["labels"].copy_(labels)
        self.model.sync_compute_weights()          # a replay runs no Python: stale bf16 copies (load_state_dict) go now
        if self.model_m is not None:
            self.model_m.sync_compute_weights()
        st["graph"].replay()
        if reducer is not None and sync and reducer.graph_events:
            reducer.exchange_after_replay()
        return st["loss"]

    def target_emb(self, input_ids):
        return self.model.decoder.get_inputs_embeds(input_ids)          # training/wrapper.py:62-63

    def compute_contrastive_loss(self, hidden_state, labels):
        """training/wrapper.py:98-118: every hidden row against the input embedding of every label of the batch (a (B*L, C) x
        (C, B*L) similarity GEMM), cross entropy against the diagonal over the non-ignored columns, weighted like the LM loss."""
        from .autograd_ops import ContrastiveLossFn, linear
        L = min(hidden_state.shape[-2], labels.shape[-1])
        labels = labels[..., :L].contiguous()
        hidden = hidden_state[..., :L, :].contiguous()
        keep = labels != self.ignore_index
        target = self.target_emb(torch.where(keep, labels, torch.zeros_like(labels)))           # (B, L, C) rows of wte
        C = hidden.shape[-1]
        cd = self.model.compute_dtype
        h2, t2 = hidden.reshape(-1, C), target.reshape(-1, C)
        pred = linear(h2 if cd == torch.float32 else h2.to(cd), t2, t2 if cd == torch.float32 else t2.to(cd), None, None, 0,
                      torch.float32)
        return ContrastiveLossFn.apply(pred, labels, L, self.contrastive_temperature, self.weight_fn, self.eos_token_weight,
                                       self.tokenizer.eos_token_id, self.ignore_index)

    def compute_lm_loss(self, lm_logits, labels, lm_logits_moco=None):
        return LmLossFn.apply(lm_logits, lm_logits_moco, labels, self.temperature, self.alpha, self.weight_fn,
                              self.eos_token_weight, self.tokenizer.eos_token_id, self.ignore_index,
                              getattr(self, "_grad_prescale", 1.0))

    def _step(self, images, labels, is_train: bool):
        tok = self.tokenizer
        keep = labels != self.ignore_index
        input_ids = torch.where(keep, labels, torch.full_like(labels, tok.eos_token_id))
        if is_train and self.mask_fraction > 0:
            mask = torch.full_like(input_ids, tok.mask_token_id)
            corrupted_mask = torch.where(torch.rand_like(input_ids, dtype=torch.float) <= self.random_mask_fraction,
                                         torch.randint_like(input_ids, low=0, high=tok.vocab_size), mask)
            corrupted = torch.where(torch.rand_like(input_ids, dtype=torch.float) <= self.mask_fraction, corrupted_mask,
                                    input_ids)
            corrupted = torch.where(keep, corrupted, torch.full_like(labels, tok.eos_token_id))
        else:
            corrupted = input_ids
        bs, sl = corrupted.shape
        bos = torch.full((bs, 1), tok.bos_token_id, device=corrupted.device, dtype=torch.long)
        corrupted = torch.cat((bos, corrupted), dim=1)[:, :sl].contiguous()
        attn_msk = torch.cat((torch.ones((bs, 1), device=keep.device, dtype=torch.bool), keep), dim=1)[:, :sl]
        step = "train" if is_train else "val"
        lm_logits, hidden_state = self(images, corrupted, attn_msk)
        lm_logits_moco = None
        if self.is_momentum and is_train:
            lm_logits_moco, _ = self.forward_m(images, corrupted, attn_msk)
        loss = self.compute_lm_loss(lm_logits, labels, lm_logits_moco)
        metrics = {f"{step}_loss_lm": loss.detach()}
        if self.add_contrastive_loss:                                   # training/wrapper.py:206-209
            loss_contrastive = self.compute_contrastive_loss(hidden_state, labels)
            metrics[f"{step}_loss_contrastive"] = loss_contrastive.detach()
            loss = loss + loss_contrastive
        if is_train:
            self._momentum_update()
        return loss, metrics
