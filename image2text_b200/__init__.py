"""image2text_b200: the encoder-decoder hot path of iitmdinesh/image2text as hand-written sm_100a CUDA behind a C ABI.

Public surface (mirrors the reference's module API):
    VisionEncoderDecoder            models/vision_encoder_decoder.py:17-182
    VisionEncoderDecoderModelOutput object_models.py:4-5
    load_training_config / TrainingConfig ...   configs/trainer.py, configs/models.py
"""
from .config_schema import (TrainingConfig, VisionEncoderDecoderConfig, load_training_config)  # noqa: F401
from .vision_encoder_decoder import VisionEncoderDecoder, VisionEncoderDecoderModelOutput  # noqa: F401

__all__ = ["VisionEncoderDecoder", "VisionEncoderDecoderModelOutput", "TrainingConfig", "VisionEncoderDecoderConfig",
           "load_training_config"]
