"""ORACLE (test infrastructure) — noise schedule + DPM-Solver++(2M) + LCM step restated from SURVEY.md
Appendix B.1-B.3 (UPSTREAM diffusers `DPMSolverMultistepScheduler` / `LCMScheduler`, as configured at
reference examples/train_flash_sdxl.py:221-236: scaled_linear betas 0.00085..0.012, 1000 steps,
dpmsolver++ order 2 midpoint, lower_order_final, final_sigmas_type="zero", timestep_spacing="trailing").

PARITY UNPINNED against diffusers itself (not installable; no golden vectors in the reference).  Written
independently of flash/schedulers.py (closed-form in float64, no shared code) so the two check each other.
"""
import math

import numpy as np
import torch


def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000, schedule="scaled_linear"):
    if schedule == "scaled_linear":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float32).astype(np.float64) ** 2
    else:
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float32).astype(np.float64)
    return np.cumprod(1.0 - betas.astype(np.float32)).astype(np.float64)


def trailing_timesteps(K, n=1000):
    return np.round(np.arange(n, 0, -n / K)).astype(np.int64) - 1


def add_noise(ac, x, eps, t):
    """sqrt(abar_t) x + sqrt(1 - abar_t) eps  (decision (1))."""
    a = torch.as_tensor(ac, dtype=x.dtype, device=x.device)[t.long()]
    shape = (-1,) + (1,) * (x.dim() - 1)
    return a.sqrt().view(shape) * x + (1 - a).sqrt().view(shape) * eps


def add_noise_schedule_index(ac, x, eps, t, K):
    """diffusers >= 0.27 `DPMSolverMultistepScheduler.add_noise` (the other reading of decision (1)):
    sigma taken BY POSITION in the current K-step trailing table; a timestep that is not on the table maps to the last
    position.  On-table timesteps give exactly the closed form."""
    ts = trailing_timesteps(K).tolist()
    sig_hat = np.sqrt((1 - ac[ts]) / ac[ts])
    idx = [ts.index(int(v)) if int(v) in ts else len(ts) - 1 for v in t.reshape(-1).tolist()]
    s = torch.as_tensor(sig_hat[idx], dtype=x.dtype, device=x.device)
    alpha_t = 1.0 / (s * s + 1.0).sqrt()
    shape = (-1,) + (1,) * (x.dim() - 1)
    return alpha_t.view(shape) * x + (s * alpha_t).view(shape) * eps


def dpm_rollout(eps_fn, x, ac, K, start_idx):
    """x_{K} from x at timesteps[start_idx] with DPM-Solver++(2M); eps_fn(x, t_int) -> eps.
    First step of the rollout and the final step are first order (Appendix B.2)."""
    ts = trailing_timesteps(K)
    sig_hat = np.sqrt((1 - ac[ts]) / ac[ts])
    sig_hat = np.concatenate([sig_hat, [0.0]])
    alpha = 1.0 / np.sqrt(sig_hat ** 2 + 1.0)
    sigma = sig_hat * alpha
    with np.errstate(divide="ignore"):
        lam = np.log(alpha) - np.log(sigma)
    x0_prev = None
    for i in range(start_idx, K):
        eps = eps_fn(x, int(ts[i]))
        x0 = (x - sigma[i] * eps) / alpha[i]
        last = i == K - 1
        if last:
            x_next = x0                                       # sigma_K = 0: alpha = 1, e^{-h} -> 0
        else:
            h = lam[i + 1] - lam[i]
            em = math.expm1(-h)
            x_next = (sigma[i + 1] / sigma[i]) * x - alpha[i + 1] * em * x0
            # order 2 (2M): second order whenever a previous x0 exists; upstream's `lower_order_second` (K < 15)
            # only demotes a third-order solver
            second = x0_prev is not None
            if second:
                r = (lam[i] - lam[i - 1]) / h
                x_next = x_next - 0.5 * alpha[i + 1] * em * (x0 - x0_prev) / r
        x0_prev = x0
        x = x_next
    return x


def lcm_scalings(t, sigma_data=0.5, timestep_scaling=10.0):
    s = t * timestep_scaling
    return sigma_data ** 2 / (s ** 2 + sigma_data ** 2), s / (s ** 2 + sigma_data ** 2) ** 0.5


def lcm_timesteps(num_steps, original_steps=50, n=1000):
    k = n // original_steps
    origin = (np.arange(1, original_steps + 1) * k - 1)[::-1]
    idx = np.floor(np.linspace(0, len(origin), num=num_steps, endpoint=False)).astype(np.int64)
    return origin[idx]
