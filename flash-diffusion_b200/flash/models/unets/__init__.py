from .unet import DiffusersUNet2DCondWrapper
from .unet2d import DiffusersUNet2DWrapper

__all__ = ["DiffusersUNet2DCondWrapper", "DiffusersUNet2DWrapper"]
