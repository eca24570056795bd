from .flash_diffusion_config import FlashDiffusionSD3Config
from .flash_diffusion_model import FlashDiffusionSD3

__all__ = ["FlashDiffusionSD3", "FlashDiffusionSD3Config"]
