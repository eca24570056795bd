"""The product's `AutoencoderKLDiffusers` wrapper logic (un-scaling, latents_mean / std, chunked encode, tiled decode
with the gaussian tile weights) against outputs of the REFERENCE's own wrapper (tests/golden/reference_vae.pt, written by
tests/golden/make_reference_vae_golden.py from the unmodified src/flash/models/vae/autoencoderKL.py + utils.py `Tiler`),
both run around the same small stand-in for the inner diffusers AutoencoderKL.  CPU: the wrapper logic is device-free."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_vae.pt"), weights_only=False)


class _Engine:
    """the stand-in behind the PRODUCT wrapper's engine interface (`decode` -> tensor, `encode_sample`)"""

    def __init__(self, fake):
        self.fake, self.config = fake, fake.config

    def decode(self, z):
        return self.fake.decode_tensor(z)

    def encode_sample(self, x, noise=None):
        return self.fake.encode_tensor(x)


@pytest.mark.parametrize("name", list(GOLD["cases"]))
def test_product_vae_wrapper_matches_reference_run(name):
    import make_reference_vae_golden as G
    from flash.models.vae import AutoencoderKLDiffusers
    rec = GOLD["cases"][name]
    case = rec["case"]
    fake = G.FakeInnerVAE(**(dict(latents_mean=G.MEAN, latents_std=G.STD) if case["stats"] else {}))
    vae = AutoencoderKLDiffusers.__new__(AutoencoderKLDiffusers)
    torch.nn.Module.__init__(vae)
    vae.__dict__["vae_model"] = _Engine(fake)
    vae.tiling_size, vae.tiling_overlap = case["tiling_size"], case["tiling_overlap"]
    vae.downsampling_factor = 2 ** (len(fake.config.block_out_channels) - 1)
    assert vae.downsampling_factor == rec["downsampling_factor"]
    vae.latent_channels = fake.config.latent_channels
    vae.latents_mean, vae.latents_std = fake.config.latents_mean, fake.config.latents_std
    vae.has_latents_mean, vae.has_latents_std = vae.latents_mean is not None, vae.latents_std is not None
    dec = vae.decode(G.latents(case))
    got, want = G.compact(dec), rec["decoded"]
    assert got["shape"] == want["shape"]
    scale = float(want["absmean"])
    assert (got["pool8"] - want["pool8"]).abs().max() < 1e-5 * max(scale, 1.0)
    assert torch.allclose(got["proj"], want["proj"], rtol=1e-4, atol=1e-3 * scale * dec.numel() ** 0.5)
    x = torch.randn(5, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    assert torch.allclose(vae.encode(x, batch_size=2), rec["encoded"], rtol=1e-6, atol=1e-7)
