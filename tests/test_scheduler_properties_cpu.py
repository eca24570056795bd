"""Size-independent PROPERTIES of the samplers (flash.schedulers), independent of diffusers and of the oracle: with an
exact denoiser for a point mass at c (x0-prediction = c, i.e. eps = (x - alpha_t c) / sigma_t), every solver used on the
path is exact — DPM-Solver++ (first order and 2M: the correction term D1 vanishes), Euler in sigma space (dx/dsigma = eps is
constant along the trajectory) and the rectified-flow Euler step (v = n - c is constant) — so any rollout, of any length
and from any start index, must stay on the ray x_t = alpha_t c + sigma_t n and end exactly at c."""
import pytest
import torch

REPO = "stabilityai/stable-diffusion-xl-base-1.0"


def _point_mass(seed, shape=(2, 4, 8, 8)):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g), torch.randn(shape, generator=g)


@pytest.mark.parametrize("K", [4, 8, 32])
@pytest.mark.parametrize("start", [0, 1, 3])
def test_dpm_solver_pp_is_exact_for_a_point_mass(K, start):
    from flash.schedulers import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler.from_pretrained(REPO, subfolder="scheduler", timestep_spacing="trailing")
    s.set_timesteps(K)
    c, n = _point_mass(K + start)
    ac = s.alphas_cumprod.double()
    ts = s.timesteps[start:]
    a0 = ac[int(ts[0])]
    x = (a0.sqrt() * c + (1 - a0).sqrt() * n).float()
    s.set_timesteps(K)                                     # fresh multistep state; the first step() finds its index from t
    for i, t in enumerate(ts):                             # a rollout from `start`, as FlashDiffusion._teacher_rollout runs it
        a = ac[int(t)]
        eps = ((x.double() - a.sqrt() * c) / (1 - a).sqrt()).float()
        assert torch.allclose(eps, n, atol=2e-3), (K, start, i)          # still on the ray of the SAME noise
        x = s.step(eps, t, x, return_dict=False)[0]
    assert torch.allclose(x, c, atol=5e-4), float((x - c).abs().max())


@pytest.mark.parametrize("K", [1, 4, 20])
def test_euler_discrete_is_exact_for_a_point_mass(K):
    from flash.schedulers import EulerDiscreteScheduler
    s = EulerDiscreteScheduler.from_pretrained(REPO, subfolder="scheduler")
    s.set_timesteps(K)
    c, n = _point_mass(K)
    sig = s.sigmas.double()
    x = (c + sig[0] * n).float()                           # on the ray x~ = c + sigma n of the sigma-space ODE
    for i, t in enumerate(s.timesteps):
        x_in = s.scale_model_input(x, t)                   # x~ / sqrt(sigma^2 + 1): what the denoiser sees
        assert torch.allclose(x_in, (x / (sig[i] ** 2 + 1).sqrt()).float(), atol=1e-6)
        eps = ((x.double() - c) / sig[i]).float()
        assert torch.allclose(eps, n, atol=2e-3)
        x = s.step(eps, t, x, return_dict=False)[0]
    assert torch.allclose(x, c, atol=5e-4), float((x - c).abs().max())


@pytest.mark.parametrize("K", [1, 4, 28])
def test_flow_match_euler_is_exact_for_a_point_mass(K):
    from flash.schedulers import FlowMatchEulerDiscreteScheduler
    s = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler")
    s.set_timesteps(K)
    c, n = _point_mass(K + 7)
    sig0 = float(s.sigmas[0])
    x = (1 - sig0) * c + sig0 * n
    for t in s.timesteps:
        x = s.step(n - c, t, x, return_dict=False)[0]      # rectified flow: v = noise - data, constant along the path
    assert torch.allclose(x, c, atol=1e-5), float((x - c).abs().max())
