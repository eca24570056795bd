"""reference: src/flash/models/flash_sd3/flash_diffusion_config.py:10-109 (same fields, defaults and
list-broadcast rules; `distill_loss_type` has no "l1" literal there)."""
from dataclasses import field
from typing import List, Literal, Optional, Union

from pydantic.dataclasses import dataclass

from ..base import ModelConfig
from ..flash.flash_diffusion_config import _per_stage


@dataclass
class FlashDiffusionSD3Config(ModelConfig):
    K: List[int] = field(default_factory=lambda: [32, 32, 32, 32, 32])
    num_iterations_per_K: List[int] = field(default_factory=lambda: [5000, 10000, 15000, 20000, 25000])
    guidance_scale_min: Union[float, List[float]] = 3.0
    guidance_scale_max: Union[float, List[float]] = 7.0
    distill_loss_type: Literal["l2", "lpips"] = "l2"
    ucg_keys: List[str] = field(default_factory=lambda: ["text"])
    timestep_distribution: Literal["gaussian", "uniform", "mixture"] = "mixture"
    mixture_num_components: Union[int, List[int]] = 4
    mixture_var: Union[float, List[float]] = 0.5
    use_dmd_loss: bool = False
    dmd_loss_scale: Union[float, List[float]] = 1.0
    distill_loss_scale: Union[float, List[float]] = 1.0
    adversarial_loss_scale: Union[float, List[float]] = 1.0
    gan_loss_type: Literal["hinge", "vanilla", "non-saturating", "wgan", "lsgan"] = "hinge"
    mode_probs: Optional[List[List[float]]] = None
    use_teacher_as_real: bool = False

    def __post_init__(self):
        super().__post_init__()
        n = len(self.K)
        self.mixture_num_components = _per_stage(self.mixture_num_components, n, int)
        for name in ("guidance_scale_min", "guidance_scale_max", "mixture_var", "distill_loss_scale",
                     "dmd_loss_scale", "adversarial_loss_scale"):
            setattr(self, name, _per_stage(getattr(self, name), n, float))
        if self.mode_probs is None:
            self.mode_probs = [[1 / m] * m for m in self.mixture_num_components]
        for i in range(n):
            assert len(self.mode_probs[i]) == self.mixture_num_components[i], (
                f"Number of mode probabilities must match number of mixture components for stage {i}, "
                f"got {len(self.mode_probs[i])} mode probabilities and {self.mixture_num_components[i]} mixture components")
        assert n == len(self.num_iterations_per_K), (
            f"Number of timesteps must match number of iterations, got {n} timesteps and "
            f"{len(self.num_iterations_per_K)} iterations")
        assert n == len(self.mode_probs), (
            f"Number of timesteps must match number of mode probabilities, got {n} timesteps and "
            f"{len(self.mode_probs)} mode probabilities")
