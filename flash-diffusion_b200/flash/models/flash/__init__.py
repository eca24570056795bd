from .flash_diffusion_config import FlashDiffusionConfig
from .flash_diffusion_model import FlashDiffusion

__all__ = ["FlashDiffusion", "FlashDiffusionConfig"]
