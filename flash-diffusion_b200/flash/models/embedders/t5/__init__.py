from .t5_embedder_model import T5TextEmbedder, T5TextEmbedderConfig

__all__ = ["T5TextEmbedder", "T5TextEmbedderConfig"]
