"""Shim of `lpips.LPIPS` (reference src/flash/models/flash/flash_diffusion_model.py:6,102-103) over flash.models.lpips."""
from flash.models.lpips import LPIPS  # noqa: F401
