"""reference: src/flash/models/embedders/timesteps/timesteps_embedding.py:6-45 (+ config :19).

The Fourier features are UPSTREAM diffusers `Timesteps` (sinusoidal, max period 10000); they are tiny host
glue executed once per batch, outside the denoiser hot path.
"""
import math

import torch
from pydantic.dataclasses import dataclass

from .base import BaseConditioner, BaseConditionerConfig


@dataclass
class TimestepsEmbedderConfig(BaseConditionerConfig):
    num_channels: int = 256
    flip_sin_to_cos: bool = True
    downscale_freq_shift: int = 0
    input_key: str = "timesteps"


def sinusoidal_embedding(t, dim, flip_sin_to_cos=True, downscale_freq_shift=0):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device)
                      / (half - downscale_freq_shift))
    ang = t.float()[:, None] * freqs[None, :]
    parts = [torch.cos(ang), torch.sin(ang)] if flip_sin_to_cos else [torch.sin(ang), torch.cos(ang)]
    emb = torch.cat(parts, dim=-1)
    if dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


class TimestepsEmbedder(BaseConditioner):
    def __init__(self, config):
        super().__init__(config)

    def forward(self, batch, force_zero_embedding: bool = False, *args, **kwargs):
        x = batch[self.input_key]
        c = self.config
        x = sinusoidal_embedding(x.flatten(), c.num_channels, c.flip_sin_to_cos,
                                 c.downscale_freq_shift).reshape(x.shape[0], -1)
        if force_zero_embedding:
            x = 0 * x
        return {self.dim2outputkey[x.dim()]: x}
