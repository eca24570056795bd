from .sd3 import DiffusersSD3Transformer2DWrapper
from .transformers import DiffusersTransformer2DWrapper

__all__ = ["DiffusersTransformer2DWrapper", "DiffusersSD3Transformer2DWrapper"]
