from .transformers import DiffusersTransformer2DWrapper

__all__ = ["DiffusersTransformer2DWrapper"]
