from .unet import DiffusersUNet2DCondWrapper


class DiffusersUNet2DWrapper:
    """reference src/flash/models/unets/unet.py:10-52: the UNCONDITIONAL `diffusers.UNet2DModel` wrapper.  No example
    script, config or test of the distillation path uses it (they all build `DiffusersUNet2DCondWrapper`), so it is not
    part of the B200 hot path; the name exists for `from flash.models.unets import ...` compatibility."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the unconditional UNet2DModel wrapper is outside the Flash-Diffusion hot path "
                                  "(SURVEY.md §8a); use DiffusersUNet2DCondWrapper")


__all__ = ["DiffusersUNet2DCondWrapper", "DiffusersUNet2DWrapper"]
