"""Shim of the `diffusers` names the Flash-Diffusion example scripts import (see ../README.md)."""
import torch

from flash.schedulers import (DDPMScheduler, DPMSolverMultistepScheduler, EulerAncestralDiscreteScheduler,  # noqa: F401
                              EulerDiscreteScheduler, FlashFlowMatchEulerDiscreteScheduler,
                              FlowMatchEulerDiscreteScheduler, LCMScheduler)

__version__ = "0.0.0+flash_b200_shim"


class _Weights:
    """stands in for `pipe.unet`: a random-init state dict in the HF key layout of the named checkpoint"""

    def __init__(self, sd):
        self._sd = sd

    def state_dict(self):
        return self._sd


def _sd15_unet_state_dict(seed=0):
    """`runwayml/stable-diffusion-v1-5` UNet keys: as DiffusersUNet2DCondWrapper's (examples/train_flash_sd.py:56-114)
    except that the Transformer2D proj_in / proj_out are 1x1 convolutions ([C, C, 1, 1]) — the script squeezes them
    (:116-158)."""
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import SD15_UNET_KWARGS, init_random_
    with torch.device("meta"):
        net = DiffusersUNet2DCondWrapper(**dict(SD15_UNET_KWARGS, use_linear_projection=True))
    net = init_random_(net.to_empty(device="cpu"), 1234 + seed)
    sd = {}
    for k, v in net.state_dict().items():
        if k.endswith(("proj_in.weight", "proj_out.weight")) and ".attentions." in k:
            v = v[:, :, None, None]
        sd[k] = v
    return sd


def _sdxl_unet_state_dict(seed=0):
    """`stabilityai/stable-diffusion-xl-base-1.0` UNet keys: the wrapper's, with the vector conditioning under
    `add_embedding.*` (the script copies it into `class_embedding`, examples/train_flash_sdxl.py:120-134)."""
    from flash.models.unets import DiffusersUNet2DCondWrapper
    from flash.recipes import SDXL_UNET_KWARGS, init_random_
    with torch.device("meta"):
        net = DiffusersUNet2DCondWrapper(**SDXL_UNET_KWARGS)
    net = init_random_(net.to_empty(device="cpu"), 1234 + seed)
    sd = {}
    for k, v in net.state_dict().items():
        sd[k.replace("class_embedding.", "add_embedding.")] = v
    return sd


class DiffusionPipeline:
    _BUILDERS = {"runwayml/stable-diffusion-v1-5": ("unet", _sd15_unet_state_dict),
                 "stabilityai/stable-diffusion-xl-base-1.0": ("unet", _sdxl_unet_state_dict)}

    @classmethod
    def from_pretrained(cls, repo, **unused):
        if repo not in cls._BUILDERS:
            raise OSError(f"{repo}: no network and no local copy; the offline shim builds random-init weights for "
                          f"{sorted(cls._BUILDERS)} only")
        attr, build = cls._BUILDERS[repo]
        pipe = cls()
        setattr(pipe, attr, _Weights(build()))
        return pipe


class StableDiffusionXLPipeline(DiffusionPipeline):
    pass


class StableDiffusion3Pipeline(DiffusionPipeline):
    pass
