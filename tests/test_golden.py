"""Golden-fixture tests (fixtures + generating script: tests/golden/).  CPU: oracle and host schedulers against the
committed vectors.  GPU: the B200 UNet engine against the same vectors (no /root/reference needed on the box)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

UNET_GOLD = torch.load(os.path.join(HERE, "golden", "tiny_unet_lora.pt"))
SCHED_GOLD = torch.load(os.path.join(HERE, "golden", "schedulers.pt"))
DIT_GOLD = torch.load(os.path.join(HERE, "golden", "tiny_dit.pt"))


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def test_oracle_unet_matches_golden():
    net = G.gold_unet()
    x, t, cond = G.gold_inputs()
    with torch.no_grad():
        out = net(x, t, cond)
        mid = net(x, t, cond, return_intermediate=True)
    assert _rel(out, UNET_GOLD["out"]) < 1e-4
    assert _rel(mid.mean(dim=(2, 3)), UNET_GOLD["mid_mean"]) < 1e-4


def test_host_schedulers_match_golden():
    from flash.schedulers import DPMSolverMultistepScheduler, LCMScheduler
    s = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    s.set_timesteps(32)
    assert torch.equal(s.timesteps, SCHED_GOLD["timesteps_K32"])
    ac = s.alphas_cumprod.double()
    assert torch.allclose(ac[[0, 499, 999]], SCHED_GOLD["alphas_cumprod_0_499_999"], rtol=1e-5)
    W = torch.tensor([[0.3, -0.2, 0.1, 0.05], [0.0, 0.25, -0.1, 0.2], [-0.15, 0.1, 0.3, 0.0], [0.2, 0.0, -0.05, 0.25]],
                     dtype=torch.float64)
    x0 = torch.linspace(-1, 1, 2 * 4 * 4 * 4, dtype=torch.float64).reshape(2, 4, 4, 4)
    for key, ref in SCHED_GOLD["rollouts"].items():
        K, start = int(key[1:key.index("_")]), int(key[key.index("_s") + 2:])
        s.set_timesteps(K)
        x = x0.clone()
        for t in s.timesteps[start:]:
            eps = torch.tanh(torch.einsum("ij,bjhw->bihw", W, x)) * (1 + int(t) / 1000.0)
            x = s.step(eps, t, x)[0]
        assert torch.allclose(x, ref, rtol=1e-4, atol=1e-5), key
    lcm = LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0")
    lcm.set_timesteps(4)
    assert torch.equal(lcm.timesteps, SCHED_GOLD["lcm_timesteps_4"])


@pytest.mark.gpu
def test_b200_unet_matches_golden():
    from flash.models.lora import LoraConfig
    from flash.models.unets import DiffusersUNet2DCondWrapper
    ora = G.gold_unet()
    net = DiffusersUNet2DCondWrapper(**G.GOLD_UNET)
    net.add_adapter(LoraConfig(r=64, lora_alpha=64, init_lora_weights="gaussian",
                               target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    net.load_state_dict(ora.state_dict())
    net = net.cuda()
    x, t, cond = G.gold_inputs()
    cond = {"cond": {k: v.cuda() for k, v in cond["cond"].items()}}
    with torch.no_grad():
        out = net(x.cuda(), t.cuda(), cond)
        mid = net(x.cuda(), t.cuda(), cond, return_intermediate=True)
    assert _rel(out.cpu(), UNET_GOLD["out"]) < 2e-2, _rel(out.cpu(), UNET_GOLD["out"])
    assert _rel(mid.mean(dim=(2, 3)).cpu(), UNET_GOLD["mid_mean"]) < 2e-2


def test_oracle_dits_and_flow_grids_match_golden():
    with torch.no_grad():
        assert _rel(G.gold_pixart()(*G.gold_dit_inputs(4, 32, 20, 64, 24, True)), DIT_GOLD["pixart_out"]) < 1e-4
        assert _rel(G.gold_sd3()(*G.gold_dit_inputs(16, 16, 9, 48, 40, False)), DIT_GOLD["sd3_out"]) < 1e-4
    from flash.schedulers import FlowMatchEulerDiscreteScheduler
    s = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    g = DIT_GOLD["flow_grids"]["train_0_499_999"]
    idx = torch.tensor([0, 499, 999])
    assert torch.equal(s.timesteps[idx], g["timesteps"]) and torch.equal(s.sigmas[idx], g["sigmas"])
    for K in (4, 32):
        s.set_timesteps(K)
        g = DIT_GOLD["flow_grids"][f"trailing_K{K}"]
        assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas, g["sigmas"])


@pytest.mark.gpu
def test_b200_dits_match_golden():
    from flash.models.transformers import DiffusersSD3Transformer2DWrapper, DiffusersTransformer2DWrapper

    def run(cls, kwargs, ora, inputs):
        with torch.device("meta"):
            net = cls(**kwargs)
        net = net.to_empty(device="cuda")
        net.load_state_dict(ora.state_dict(), strict=False)
        if hasattr(ora.pos_embed, "pos_embed"):
            net.pos_embed.pos_embed = ora.pos_embed.pos_embed.clone().cuda()
        net.freeze()
        x, t, cond = inputs
        cond = {"cond": {k: v.cuda() for k, v in cond["cond"].items()}}
        with torch.no_grad():
            return net(x.cuda(), t.cuda(), cond).cpu()
    px = run(DiffusersTransformer2DWrapper, G.GOLD_PIXART, G.gold_pixart(), G.gold_dit_inputs(4, 32, 20, 64, 24, True))
    assert _rel(px, DIT_GOLD["pixart_out"]) < 2e-2, _rel(px, DIT_GOLD["pixart_out"])
    s3 = run(DiffusersSD3Transformer2DWrapper, G.GOLD_SD3, G.gold_sd3(), G.gold_dit_inputs(16, 16, 9, 48, 40, False))
    assert _rel(s3, DIT_GOLD["sd3_out"]) < 2e-2, _rel(s3, DIT_GOLD["sd3_out"])
