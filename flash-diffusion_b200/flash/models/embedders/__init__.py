"""Conditioner abstraction (reference: src/flash/models/embedders/__init__.py:1-21, same exported names).

The CLIP / T5 text conditioners keep the reference's classes and contracts; their encoders are `transformers` library
models on the far side of the hot-path boundary, built offline with random weights
(offline_text.py).  The benchmarks feed synthetic embeddings through `TorchNNEmbedder` / `TimestepsEmbedder`
(SURVEY.md Appendix C recipe) or `PrecomputedTextEmbedder`.
"""
from .base import BaseConditioner, BaseConditionerConfig
from .clip import ClipEmbedder, ClipEmbedderConfig, ClipEmbedderWithProjection
from .conditioners_wrapper import ConditionerWrapper
from .precomputed import PrecomputedTextEmbedder, PrecomputedTextEmbedderConfig
from .t5 import T5TextEmbedder, T5TextEmbedderConfig
from .timesteps import TimestepsEmbedder, TimestepsEmbedderConfig
from .torch_nn import TorchNNEmbedder, TorchNNEmbedderConfig

__all__ = ["BaseConditioner", "BaseConditionerConfig", "ClipEmbedder", "ClipEmbedderConfig", "ClipEmbedderWithProjection",
           "ConditionerWrapper", "PrecomputedTextEmbedder", "PrecomputedTextEmbedderConfig", "T5TextEmbedder",
           "T5TextEmbedderConfig", "TimestepsEmbedder", "TimestepsEmbedderConfig", "TorchNNEmbedder",
           "TorchNNEmbedderConfig"]
