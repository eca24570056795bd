"""ORACLE (test infrastructure — never imported by the product path).

Pure-PyTorch fp32 restatement of the UPSTREAM `diffusers.models.AutoencoderKL` behind the reference's VAE wrapper
`AutoencoderKLDiffusers` (reference src/flash/models/vae/autoencoderKL.py:9-128: `encode` = posterior sample times
`scaling_factor` (:52-62), `decode` = (z / scaling_factor) through the decoder (:64-128)), with the architecture of the
checkpoints the example scripts name (`runwayml/stable-diffusion-v1-5` subfolder "vae", `stabilityai/sdxl-vae`):
block_out_channels (128, 256, 512, 512), layers_per_block 2, 32 groups, eps 1e-6, single-head mid attention, 4 latent
channels, state-dict keys as in diffusers (`encoder.down_blocks.0.resnets.0.norm1.weight`, `quant_conv.weight`, ...).

PARITY UNPINNED: diffusers is not installable offline and the reference holds no golden vectors; the
file is pinned structurally (parameter count 83,653,863 of the SD / SDXL VAE, key scheme) and by the reference's own
shape tests (tests/test_vaes/test_autoencoderKL.py:31-44: 32x32 px <-> 4x4 latents, factor 8).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

VAE_KWARGS = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, norm_num_groups=32)
SCALING = {"runwayml/stable-diffusion-v1-5": 0.18215, "stabilityai/sdxl-vae": 0.13025,
           "stabilityai/stable-diffusion-xl-base-1.0": 0.13025}


class ResnetBlock(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, bias=True, norm_num_groups, residual_connection=True, eps 1e-6)."""

    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        a = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
        o = self.to_out[0](a @ v)
        return o.transpose(1, 2).reshape(B, C, H, W) + x


class MidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(ch, ch, groups), ResnetBlock(ch, ch, groups)])
        self.attentions = nn.ModuleList([Attention(ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Downsample(nn.Module):
    """diffusers Downsample2D(padding=0): pad right/bottom by one, 3x3 stride-2 conv without padding."""

    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class Upsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.downsamplers is None else self.downsamplers[0](x)


class UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(DownEncoderBlock(c, co, layers, groups, i != len(boc) - 1))
            c = co
        self.mid_block = MidBlock(c, groups)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, out_channels, latent_channels, boc, layers, groups):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = MidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(UpDecoderBlock(c, co, layers + 1, groups, i != len(rev) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLOracle(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, use_quant_conv=True,
                 use_post_quant_conv=True):
        super().__init__()
        boc = list(block_out_channels)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.decoder = Decoder(out_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        # the SD3 VAE config switches both 1x1 convs off (identity)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1) if use_quant_conv else nn.Identity()
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1) if use_post_quant_conv else nn.Identity()
        self.scaling_factor = scaling_factor
        self.latent_channels = latent_channels

    def moments(self, x):
        """(mean, logvar) of the diagonal Gaussian posterior (logvar clamped to [-30, 20] as upstream)."""
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)

    def encode(self, x, noise=None):
        """reference wrapper `encode` (:52-62): latent_dist.sample() * scaling_factor; `noise` makes the draw explicit."""
        mean, logvar = self.moments(x)
        if noise is None:
            noise = torch.randn_like(mean)
        return (mean + torch.exp(0.5 * logvar) * noise) * self.scaling_factor

    def decode(self, z):
        """reference wrapper `decode` (:64-77,:126): decoder(post_quant_conv(z / scaling_factor))."""
        return self.decoder(self.post_quant_conv(z / self.scaling_factor))
