"""Generates the committed golden fixtures (run in the build container, CPU only):

    python tests/golden/make_golden.py

The reference itself cannot run here (diffusers fork / peft / lightning are not installable) and
holds no golden vectors, so these fixtures are ORACLE-generated: they pin the oracle (and, through the GPU tests,
the kernels) against regressions and travel to the GPU box where /root/reference and the build container's CPU
state do not exist.  Parity with the upstream libraries stays "unpinned" (see oracle/unet.py header).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import flash_step_sd3 as O3  # noqa: E402
from oracle import schedulers as OS  # noqa: E402
from oracle.dit import PixArtTransformerOracle  # noqa: E402
from oracle.sd3 import SD3TransformerOracle  # noqa: E402
from oracle.unet import LoraConfig, UNet2DConditionOracle  # noqa: E402

GOLD_UNET = dict(in_channels=4, out_channels=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"],
                 up_block_types=["CrossAttnUpBlock2D", "UpBlock2D"], block_out_channels=[64, 128], layers_per_block=1,
                 cross_attention_dim=96, transformer_layers_per_block=[1, 2], attention_head_dim=[1, 2],
                 use_linear_projection=True, class_embed_type="projection", projection_class_embeddings_input_dim=48)


def seeded_state_dict(net, seed):
    """Weights as a pure function of (parameter name order, seed): regenerated on the GPU box, not stored."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, p in net.state_dict().items():
        if p.dim() >= 2:
            sd[name] = torch.randn(p.shape, generator=g) / (p[0].numel() ** 0.5)
        elif "lora" in name:
            sd[name] = torch.randn(p.shape, generator=g) * 0.05
        elif name.endswith("bias"):
            sd[name] = torch.randn(p.shape, generator=g) * 0.05
        else:
            sd[name] = 1.0 + torch.randn(p.shape, generator=g) * 0.05
    for name in sd:
        if "lora_B" in name:
            sd[name] = torch.randn(sd[name].shape, generator=g) * 0.05
    return sd


def gold_unet(lora=True):
    net = UNet2DConditionOracle(**GOLD_UNET)
    if lora:
        net.add_adapter(LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian",
                                   target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    net.load_state_dict(seeded_state_dict(net, 2024))
    return net


def gold_inputs():
    g = torch.Generator().manual_seed(99)
    x = torch.randn(2, 4, 32, 32, generator=g)
    t = torch.tensor([999.0, 250.0])
    cond = {"cond": {"crossattn": torch.randn(2, 77, 96, generator=g), "vector": torch.randn(2, 48, generator=g)}}
    return x, t, cond


GOLD_PIXART = dict(sample_size=32, num_layers=2, attention_head_dim=24, in_channels=4, out_channels=8, patch_size=2,
                   attention_bias=True, num_attention_heads=4, cross_attention_dim=96, activation_fn="gelu-approximate",
                   norm_type="ada_norm_single", norm_elementwise_affine=False, norm_eps=1e-6, caption_channels=64,
                   projection_class_embeddings_input_dim=8, time_embed_dim=96, timesteps_embedding_num_channels=32,
                   use_concat_vector_conditioning=True, num_vector_conditionings=3)
GOLD_SD3 = dict(sample_size=16, patch_size=2, in_channels=16, num_layers=3, attention_head_dim=64, num_attention_heads=2,
                joint_attention_dim=48, caption_projection_dim=128, pooled_projection_dim=40, out_channels=16,
                pos_embed_max_size=12)


def _seeded(net, seed):
    sd = seeded_state_dict(net, seed)
    for name, buf in net.named_buffers():          # deterministic position tables stay as constructed
        if name in sd:
            sd[name] = buf.clone()
    net.load_state_dict(sd)
    return net


def gold_pixart():
    return _seeded(PixArtTransformerOracle(**GOLD_PIXART), 2025)


def gold_sd3():
    return _seeded(SD3TransformerOracle(**GOLD_SD3), 2026)


def gold_dit_inputs(channels, hw, tokens, ctx_dim, vec_dim, masked):
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, channels, hw, hw, generator=g)
    t = torch.tensor([999.0, 250.0])
    cond = {"cond": {"crossattn": torch.randn(2, tokens, ctx_dim, generator=g), "vector": torch.randn(2, vec_dim, generator=g)}}
    if masked:
        cond["cond"]["attention_mask"] = (torch.arange(tokens)[None, :] < torch.tensor([[tokens - 6], [tokens - 3]])).long()
    return x, t, cond


def main():
    with torch.no_grad():
        px = gold_pixart()(*gold_dit_inputs(4, 32, 20, 64, 24, True))
        s3 = gold_sd3()(*gold_dit_inputs(16, 16, 9, 48, 40, False))
    grids = {}
    for K in (4, 32):
        ts, sig = O3.inference_grid(K)
        grids[f"trailing_K{K}"] = {"timesteps": ts, "sigmas": sig}
    tt, ss = O3.training_grid()
    grids["train_0_499_999"] = {"timesteps": tt[[0, 499, 999]], "sigmas": ss[[0, 499, 999]]}
    torch.save({"pixart_out": px, "sd3_out": s3, "flow_grids": grids}, os.path.join(HERE, "tiny_dit.pt"))
    net = gold_unet()
    x, t, cond = gold_inputs()
    with torch.no_grad():
        out = net(x, t, cond)
        mid = net(x, t, cond, return_intermediate=True)
    torch.save({"out": out, "mid_mean": mid.mean(dim=(2, 3)), "mid_abs_mean": mid.abs().mean()},
               os.path.join(HERE, "tiny_unet_lora.pt"))
    ac = OS.alphas_cumprod()
    W = torch.tensor([[0.3, -0.2, 0.1, 0.05], [0.0, 0.25, -0.1, 0.2], [-0.15, 0.1, 0.3, 0.0], [0.2, 0.0, -0.05, 0.25]],
                     dtype=torch.float64)
    x0 = torch.linspace(-1, 1, 2 * 4 * 4 * 4, dtype=torch.float64).reshape(2, 4, 4, 4)

    def eps_fn(x, t):
        return torch.tanh(torch.einsum("ij,bjhw->bihw", W, x)) * (1 + t / 1000.0)

    traj = {f"K{K}_s{s}": OS.dpm_rollout(eps_fn, x0.clone(), ac, K, s) for K, s in [(32, 0), (32, 8), (32, 16), (32, 24), (4, 1)]}
    torch.save({"timesteps_K32": torch.from_numpy(OS.trailing_timesteps(32)), "rollouts": traj,
                "alphas_cumprod_0_499_999": torch.tensor([ac[0], ac[499], ac[999]], dtype=torch.float64),
                "lcm_timesteps_4": torch.from_numpy(OS.lcm_timesteps(4).copy())},
               os.path.join(HERE, "schedulers.pt"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
