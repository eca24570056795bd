from .autoencoderKL import AutoencoderKL, AutoencoderKLDiffusers
from .autoencoderKL_config import AutoencoderKLDiffusersConfig

__all__ = ["AutoencoderKL", "AutoencoderKLDiffusers", "AutoencoderKLDiffusersConfig"]
