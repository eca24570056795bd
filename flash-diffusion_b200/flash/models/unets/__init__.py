from .unet import DiffusersUNet2DCondWrapper

__all__ = ["DiffusersUNet2DCondWrapper"]
