"""Golden vectors from the REFERENCE's own VAE wrapper (src/flash/models/vae/autoencoderKL.py:9-128: un-scaling,
latents_mean / std, chunked encode, the tiled decode through `Tiler.get_tiles` / `pad` / `Tiler.merge_tiles` with its
gaussian tile weights, src/flash/models/utils.py:12-262,333-349), imported unmodified from /root/reference/src:
    python tests/golden/make_reference_vae_golden.py  ->  tests/golden/reference_vae.pt
The inner `diffusers.models.AutoencoderKL` (not installable) is a small deterministic stand-in with a NON-pointwise
decoder (3x3 box filter after the x8 upsampling), so that how overlapping tiles are weighted shows in the result; the
product's `AutoencoderKLDiffusers` wrapper is then run around the same stand-in (tests/test_reference_vae_golden.py)."""
import os
import sys
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)


class FakeInnerVAE:
    """the arithmetic both wrappers are run around"""
    config = types.SimpleNamespace(in_channels=3, out_channels=3, latent_channels=4, scaling_factor=0.5,
                                   latents_mean=None, latents_std=None, block_out_channels=[8, 8, 8, 8])

    def __init__(self, latents_mean=None, latents_std=None):
        self.config = types.SimpleNamespace(**vars(FakeInnerVAE.config))
        self.config.latents_mean, self.config.latents_std = latents_mean, latents_std

    @staticmethod
    def encode_tensor(x):
        p = F.avg_pool2d(x, 8)
        return torch.cat([p, p.mean(1, keepdim=True)], dim=1)

    @staticmethod
    def decode_tensor(z):
        up = F.interpolate(z[:, :3] + 0.1 * z[:, 3:4], scale_factor=8, mode="nearest")
        k = torch.ones(3, 1, 3, 3) / 9.0
        return F.conv2d(F.pad(up, (1, 1, 1, 1), mode="replicate"), k, groups=3) + 0.01 * up ** 2

    # diffusers-shaped API (what the reference wrapper calls)
    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls(**getattr(cls, "_next_kwargs", {}))

    def encode(self, x):
        z = self.encode_tensor(x)
        return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda: z))

    def decode(self, z):
        return types.SimpleNamespace(sample=self.decode_tensor(z))


CASES = [dict(name="untiled_32", shape=(2, 4, 32, 32), tiling_size=(64, 64), tiling_overlap=(16, 16), stats=False),
         dict(name="tiled_96x80", shape=(1, 4, 96, 80), tiling_size=(64, 64), tiling_overlap=(16, 16), stats=False),
         dict(name="tiled_rows_100x64_b2", shape=(2, 4, 100, 64), tiling_size=(64, 64), tiling_overlap=(16, 16), stats=False),
         dict(name="tiled_40x72_small_tiles_stats", shape=(1, 4, 40, 72), tiling_size=(32, 32), tiling_overlap=(8, 8), stats=True)]
MEAN, STD = [0.1, -0.2, 0.05, 0.3], [1.5, 0.7, 1.1, 0.9]


def latents(case):
    g = torch.Generator().manual_seed(sum(case["shape"]))
    return torch.randn(*case["shape"], generator=g)


def compact(dec):
    """what is stored / compared instead of the full-resolution image: its 8x8 block means and eight seeded random
    projections of the full-resolution tensor (which see every pixel)"""
    dec = dec.float()
    g = torch.Generator().manual_seed(123)
    proj = torch.stack([(dec * torch.randn(dec.shape, generator=g)).sum() for _ in range(8)])
    return dict(pool8=F.avg_pool2d(dec, 8).clone(), proj=proj, shape=tuple(dec.shape), absmean=dec.abs().mean())


def main():
    import make_reference_step_golden as G
    G.install_shims()
    sys.modules["diffusers.models"].AutoencoderKL = FakeInnerVAE
    sys.path.insert(0, G.REF_SRC)
    from flash.models.vae import AutoencoderKLDiffusers, AutoencoderKLDiffusersConfig
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    out = dict(cases={}, generated_by=os.path.relpath(__file__, ROOT),
               reference_files=["src/flash/models/vae/autoencoderKL.py", "src/flash/models/utils.py"])
    for case in CASES:
        FakeInnerVAE._next_kwargs = dict(latents_mean=MEAN, latents_std=STD) if case["stats"] else {}
        vae = AutoencoderKLDiffusers(AutoencoderKLDiffusersConfig(
            version="stand-in", tiling_size=case["tiling_size"], tiling_overlap=case["tiling_overlap"]))
        z = latents(case)
        dec = vae.decode(z.clone())
        x = torch.randn(5, 3, 64, 64, generator=torch.Generator().manual_seed(9))
        enc = vae.encode(x, batch_size=2)
        out["cases"][case["name"]] = dict(case=case, decoded=compact(dec), encoded=enc.clone(),
                                          downsampling_factor=vae.downsampling_factor)
        print(case["name"], tuple(dec.shape), "factor", vae.downsampling_factor)
    path = os.path.join(HERE, "reference_vae.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
