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
    s = DPMSolverMultistepScheduler.from_pretrained("x", timestep_spacing="trailing")
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
    lcm = LCMScheduler.from_pretrained("x")
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
