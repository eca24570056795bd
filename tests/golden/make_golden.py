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

from oracle import schedulers as OS  # noqa: E402
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


def main():
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
