from .clip_embedder_config import ClipEmbedderConfig
from .clip_embedder_model import ClipEmbedder, ClipEmbedderWithProjection

__all__ = ["ClipEmbedder", "ClipEmbedderWithProjection", "ClipEmbedderConfig"]
