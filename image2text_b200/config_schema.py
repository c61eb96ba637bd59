"""YAML surface of the reference, re-declared (same field names, so the reference's
``training_configs/**.yaml`` files parse unchanged).

Mirrors ``configs/models.py:9-135`` and ``configs/trainer.py:6-41`` of the reference.  Only the
schema is shared; which fields the B200 path implements is stated in DESIGN.md (unsupported
combinations raise ``NotImplementedError`` at model construction, they are never silently
ignored).
"""
from __future__ import annotations

from enum import Enum
from typing import List, Optional, Tuple, Union

import yaml
from pydantic import BaseModel


class LoraSpec(BaseModel):
    r: int = 16
    lora_alpha: int = 64
    lora_dropout: float = 0.1
    target_modules: Optional[List[str]] = None
    force_enable_update_modules: Optional[List[str]] = None


class MLPConfig(BaseModel):
    ff_mult: float


class MoEConfig(BaseModel):
    num_experts: int
    proj_features: int
    ff_mult_factor: float
    gate_sizes: Optional[Tuple[int, ...]] = None
    top_k: int = 1


class SelfAttentionType(Enum):
    MULTI_HEAD = "multi_head"
    MULTI_QUERY = "multi_query"


class SelfAttentionConfig(BaseModel):
    attn_type: SelfAttentionType
    n_embd: int = 768
    n_head: int = 12
    bias: bool = True
    dropout: float = 0.1
    attn_dropout: float = 0.1


class TransformerConfig(BaseModel):
    attn_config: SelfAttentionConfig
    rotator_config: Union[MoEConfig, MLPConfig]
    is_causal: bool = False
    is_cross_attn: bool = False
    is_sparse_attn: bool = False
    sparsity_factor: float = 0.5
    max_block_size: Optional[int] = None


class ImageInputSpec(BaseModel):
    width: int
    height: int
    n_channels: int = 3


class LshConfig(BaseModel):
    num_bins: Tuple[int, ...]
    num_proj: int
    learnable: bool


class PeerConfig(BaseModel):
    num_units_sqrt: int
    topk: int
    nhead: int
    query_dim: Optional[int] = None


class EncoderConfig(BaseModel):
    n_cls: int
    lora_spec: Optional[LoraSpec] = None


class VisionTransformerEncoderConfig(EncoderConfig):
    transformer_config: TransformerConfig
    input: ImageInputSpec
    num_patches: int
    n_channels: int
    n_layer: int = 12
    enable_gradient_checkpointing: bool = False
    feature_extractor_gate_sizes: Optional[Tuple[int, ...]] = None
    feature_extractor_kernel_size: Tuple[int, int] = (4, 4)


class PretrainedViTConfig(EncoderConfig):
    n_embd_out_vit: int
    refine_base_model: bool = True
    gate_sizes: Optional[Tuple[int, ...]] = None
    lsh_config: Optional[LshConfig] = None
    peer_config: Optional[PeerConfig] = None


class ModelType(Enum):
    GPT2 = "gpt2"
    GPT2_MEDIUM = "gpt2-medium"
    GPT2_LARGE = "gpt2-large"
    GPT2_XL = "gpt2-xl"


class DecoderConfig(BaseModel):
    vocab_size: int
    lora_spec: Optional[LoraSpec] = None
    enable_gradient_checkpointing: bool = False


class TransformerDecoderConfig(DecoderConfig):
    transformer_config: TransformerConfig
    n_layer: int
    block_size: int
    pretrained_model: Optional[ModelType] = None
    skip_alternate_cross_attn: bool = True
    use_advanced_pos_emb: bool = False
    advanced_pos_emb_gate_sizes: Optional[Tuple[int, ...]] = None


class HuggingfaceDecoderConfig(DecoderConfig):
    model_str: str
    use_cross_attn: bool
    extra_tokens: int
    load_in_4bit: bool
    prepare_for_kbit_training: bool
    use_auth_token: bool = False


class VisionEncoderDecoderConfig(BaseModel):
    vision_encoder_config: Union[VisionTransformerEncoderConfig, PretrainedViTConfig]
    decoder_config: Union[TransformerDecoderConfig, HuggingfaceDecoderConfig]
    use_cross_attn: bool = False
    use_soft_prompting: bool = True
    no_repeat_n_grams: Tuple[int, ...] = (2, 3, 4, 5)
    loose_match_decoder_state_dict: bool = False
    chkpt_path: Optional[str] = None


class TrainerWrapperConfig(BaseModel):
    moco_momentum: Optional[float] = None
    moco_alpha: Optional[float] = None
    training_temperature: float = 1.0
    weight_fn: str = "constant"
    mask_fraction: float = 0.0
    random_mask_fraction: float = 0.0
    eos_token_weight: Optional[float] = None
    add_contrastive_loss: bool = False
    training_contrastive_temperature: float = 1.0


class OptimizerConfig(BaseModel):
    lr: float
    weight_decay: float = 0.0
    betas: Tuple[float, float] = (0.9, 0.999)
    target_modules: Optional[List[str]] = None


class TrainingConfig(BaseModel):
    model: VisionEncoderDecoderConfig
    trainer: TrainerWrapperConfig
    optimizers: List[OptimizerConfig]
    tokenizer_str: str
    batch_size: int
    gradient_accumulation_steps: int = 1
    precision: str = "no"
    epochs: int = 1
    num_steps: Optional[int] = None
    num_val_steps: Optional[int] = None
    ignore_index: int = -100
    shuffle: bool = True
    dataloader_buffer_size: int = 5
    disable_flash: bool = False
    reset_moco_after_k_epochs: Optional[List[int]] = None
    use_snr_optim: bool = False


def load_training_config(path: str) -> TrainingConfig:
    """Same entry as ``trainer.py:106-107``: yaml.safe_load -> TrainingConfig."""
    with open(path, "r") as fh:
        return TrainingConfig.model_validate(yaml.safe_load(fh))
