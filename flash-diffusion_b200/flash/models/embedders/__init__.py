"""Conditioner abstraction (reference: src/flash/models/embedders/__init__.py:1-21).

CLIP / T5 text encoders are out of scope (frozen, need HF weights): the hot path is
fed synthetic embeddings through `TorchNNEmbedder` / `TimestepsEmbedder` (SURVEY.md Appendix C recipe).
"""
from .base import BaseConditioner, BaseConditionerConfig
from .conditioners_wrapper import ConditionerWrapper
from .precomputed import PrecomputedTextEmbedder, PrecomputedTextEmbedderConfig
from .timesteps import TimestepsEmbedder, TimestepsEmbedderConfig
from .torch_nn import TorchNNEmbedder, TorchNNEmbedderConfig

__all__ = ["PrecomputedTextEmbedder", "PrecomputedTextEmbedderConfig", "BaseConditioner", "BaseConditionerConfig", "ConditionerWrapper", "TimestepsEmbedder",
           "TimestepsEmbedderConfig", "TorchNNEmbedder", "TorchNNEmbedderConfig"]
