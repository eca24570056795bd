"""Noise schedulers consumed by FlashDiffusion (protocol: SURVEY.md Appendix A "Scheduler protocol").

The reference takes these from diffusers (`DDPMScheduler`, `DPMSolverMultistepScheduler`, `LCMScheduler`;
reference call sites src/flash/models/flash/flash_diffusion_model.py:139,172,245-257,289-324,781-863 and
examples/train_flash_sdxl.py:221-236).  diffusers is not installable here, so the scheduler math is
restated from its published algorithms; all of it is scalar host arithmetic
plus one elementwise update per step — on CUDA that update runs in the fused kernel
`fd_step_cfg_dpm` (CFG combine + DPM-Solver++ step), see `DPMSolverMultistepScheduler.fused_cfg_step`.

Decision (1) of SURVEY.md §8c: `add_noise` is the closed form sqrt(abar_t) x + sqrt(1-abar_t) eps for any
integer timestep (off-schedule DMD / GAN timesteps included).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def _betas(schedule: str, beta_start: float, beta_end: float, n: int) -> torch.Tensor:
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(schedule)


def _spaced_timesteps(num_train: int, num_inference: int, spacing: str, steps_offset: int = 0) -> np.ndarray:
    if spacing == "trailing":
        return np.round(np.arange(num_train, 0, -num_train / num_inference)).astype(np.int64) - 1
    if spacing == "leading":
        ratio = num_train // num_inference
        return (np.arange(0, num_inference) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset
    if spacing == "linspace":
        return np.linspace(0, num_train - 1, num_inference).round()[::-1].copy().astype(np.int64)
    raise NotImplementedError(spacing)


# repo -> scheduler config used by `from_pretrained` (there is no network; these are the published
# scheduler_config.json values of the checkpoints the reference examples name)
_SD_CFG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
               steps_offset=1)
_KNOWN_CONFIGS = {
    # examples/train_flash_sdxl.py:221-236 and examples/train_flash_sd.py:204-219 (SD1.5 also reads the SDXL repo)
    "stabilityai/stable-diffusion-xl-base-1.0": _SD_CFG,
    "runwayml/stable-diffusion-v1-5": _SD_CFG,
    # examples/train_flash_pixart.py:259-274: the published PixArt-alpha scheduler_config.json is a
    # DPMSolverMultistepScheduler over LINEAR betas 1e-4 .. 0.02, steps_offset 0 (restated from memory: the file is
    # not available offline — flagged "parity unpinned" like every upstream constant)
    "PixArt-alpha/PixArt-XL-2-1024-MS": dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                                             beta_schedule="linear", steps_offset=0),
}



def _host_tables(cls):
    """Scheduler tables are a few thousand floats that live on the host whatever the ambient default device is (the
    full-size example scripts are plumbing-tested under `torch.device("meta")`): run the table-building methods of
    `cls` under an explicit CPU device context."""
    import functools
    for name in ("__init__", "set_timesteps"):
        fn = cls.__dict__.get(name)
        if fn is None:
            continue

        def make(fn):
            @functools.wraps(fn)
            def wrapped(*a, **k):
                if torch.get_default_device().type == "cpu":
                    return fn(*a, **k)
                with torch.device("cpu"):
                    return fn(*a, **k)
            return wrapped
        setattr(cls, name, make(fn))
    return cls


@_host_tables
class _SchedulerBase:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 timestep_spacing="leading", steps_offset=0, prediction_type="epsilon", **extra):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset,
                                      prediction_type=prediction_type, **extra)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    @classmethod
    def from_pretrained(cls, repo: str = None, subfolder: str = None, revision: str = None, **overrides):
        """There is no network: `repo` selects one of the scheduler configs of the checkpoints the reference example
        scripts name; an unknown repo raises instead of silently training on the wrong alphas_cumprod."""
        if repo not in _KNOWN_CONFIGS:
            raise ValueError(f"unknown scheduler repo {repo!r}: offline build, known configs: {sorted(_KNOWN_CONFIGS)}")
        cfg = dict(_KNOWN_CONFIGS[repo])
        cfg.update(overrides)
        return cls(**cfg)

    @classmethod
    def from_config(cls, config, **overrides):
        cfg = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        cfg.update(overrides)
        return cls(**cfg)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = timesteps.to(original_samples.device).long()
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return ac[t].sqrt().view(shape) * original_samples + (1 - ac[t]).sqrt().view(shape) * noise

    def _default_set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        ts = _spaced_timesteps(self.config.num_train_timesteps, num_inference_steps,
                               self.config.timestep_spacing, self.config.steps_offset)
        self.timesteps = torch.from_numpy(ts)


@_host_tables
class DDPMScheduler(_SchedulerBase):
    """Ancestral DDPM sampler (epsilon prediction, fixed_small variance, clip_sample as in diffusers'
    defaults).  The reference's own test builds every scheduler as `DDPMScheduler()`
    (tests/test_flash/test_flash_diffusion.py:87-93)."""

    def __init__(self, clip_sample=True, clip_sample_range=1.0, variance_type="fixed_small", **kw):
        super().__init__(clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                         variance_type=variance_type, **kw)

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            self.timesteps = torch.as_tensor(np.array(timesteps, dtype=np.int64))
            self.num_inference_steps = len(self.timesteps)
            return
        self._default_set_timesteps(num_inference_steps)

    def _prev_timestep(self, t):
        n = self.num_inference_steps or self.config.num_train_timesteps
        return t - self.config.num_train_timesteps // n

    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        t = int(timestep)
        prev_t = self._prev_timestep(t)
        ac_t = float(self.alphas_cumprod[t])
        ac_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else 1.0
        beta_prod_t, beta_prod_prev = 1 - ac_t, 1 - ac_prev
        cur_alpha = ac_t / ac_prev
        cur_beta = 1 - cur_alpha
        x0 = (sample - beta_prod_t ** 0.5 * model_output) / ac_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        c0 = (ac_prev ** 0.5 * cur_beta) / beta_prod_t
        ct = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            var = max(beta_prod_prev / beta_prod_t * cur_beta, 1e-20)
            prev = prev + var ** 0.5 * torch.randn(sample.shape, generator=generator, device=sample.device,
                                                   dtype=sample.dtype)
        return (prev,)


@_host_tables
class DPMSolverMultistepScheduler(_SchedulerBase):
    """DPM-Solver++ (2M, midpoint), epsilon prediction, `final_sigmas_type="zero"`, `lower_order_final`
    (SURVEY.md Appendix B.2).  `set_timesteps` resets the multistep history, so a rollout entered at
    `timesteps[start_idx:]` starts first-order, exactly as in the reference loop
    (flash_diffusion_model.py:139,288-324)."""

    order = 1

    def __init__(self, solver_order=2, timestep_spacing="leading", add_noise_mode="closed_form", **kw):
        """add_noise_mode (SURVEY.md §8c decision 1; the reference's diffusers fork is a moving branch, so BOTH readings
        of `add_noise` for the off-schedule DMD / GAN timesteps (flash_diffusion_model.py:416-428, :524-539) exist):
          "closed_form"     sqrt(abar_t) x + sqrt(1 - abar_t) eps for any integer t (diffusers < 0.26; default, and what
                            the model's own `sqrt_alpha_cumprod` / `sigmas` buffers at :110-119 assume);
          "schedule_index"  diffusers >= 0.27: sigma looked up BY POSITION in the current K-step table
                            (`index_for_timestep`), a timestep that is not on the table maps to the LAST index."""
        kw.pop("algorithm_type", None)
        if add_noise_mode not in ("closed_form", "schedule_index"):
            raise ValueError(f"add_noise_mode={add_noise_mode!r}")
        if solver_order > 2:
            raise NotImplementedError("third-order DPM-Solver++ update (no example script configures solver_order=3)")
        self.add_noise_mode = add_noise_mode
        super().__init__(timestep_spacing=timestep_spacing, solver_order=solver_order, **kw)
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self._step_index = None
        self.set_timesteps(self.config.num_train_timesteps)

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            ts = np.array(timesteps, dtype=np.int64)
            self.num_inference_steps = len(ts)
        else:
            ts = _spaced_timesteps(self.config.num_train_timesteps, num_inference_steps,
                                   self.config.timestep_spacing, self.config.steps_offset)
            ts = np.clip(ts, 0, self.config.num_train_timesteps - 1)
            self.num_inference_steps = num_inference_steps
        ac = self.alphas_cumprod.double().numpy()
        sig = np.sqrt((1 - ac) / ac)[ts]
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]))          # final sigma = 0 (float64)
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self._step_index = None

    def add_noise(self, original_samples, noise, timesteps):
        if self.add_noise_mode == "closed_form":
            return super().add_noise(original_samples, noise, timesteps)
        sched = self.timesteps.tolist()
        last = len(sched) - 1
        pos = {}
        for i, t in enumerate(sched):          # second occurrence wins when a timestep is duplicated (upstream)
            pos.setdefault(t, []).append(i)
        idx = [(pos[t][1] if len(pos[t]) > 1 else pos[t][0]) if t in pos else last
               for t in timesteps.reshape(-1).tolist()]
        sig = self.sigmas[idx].to(device=original_samples.device, dtype=original_samples.dtype)
        alpha_t = 1.0 / (sig * sig + 1.0).sqrt()
        sigma_t = sig * alpha_t
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return alpha_t.view(shape) * original_samples + sigma_t.view(shape) * noise

    @staticmethod
    def _alpha_sigma(sigma: float):
        alpha_t = 1.0 / math.sqrt(sigma * sigma + 1.0)
        return alpha_t, sigma * alpha_t

    def _index_for(self, timestep) -> int:
        t = int(timestep)
        idx = (self.timesteps == t).nonzero()
        if len(idx) == 0:
            return len(self.timesteps) - 1
        return int(idx[1] if len(idx) > 1 else idx[0])

    def step_coefficients(self, timestep):
        """Host scalars of one update at `timestep`:
           x0 = (x - sigma_t eps) / alpha_t ;  x_next = c_x x - c_d0 x0 - c_d1r (x0 - x0_prev).
        Returns (alpha_t, sigma_t, c_x, c_d0, c_d1r) and advances the multistep state."""
        if self._step_index is None:
            self._step_index = self._index_for(timestep)
        i = self._step_index
        n = len(self.timesteps)
        sig_s0, sig_t = float(self.sigmas[i]), float(self.sigmas[i + 1])
        a_s0, s_s0 = self._alpha_sigma(sig_s0)
        a_t, s_t = self._alpha_sigma(sig_t)
        lam_s0 = math.log(a_s0) - math.log(s_s0)
        # upstream order selection: `order==1 or lower_order_nums<1 or lower_order_final` -> 1st; `order==2 or ...`
        # -> 2nd.  `lower_order_second` only demotes a THIRD-order solver, so it plays no role at solver_order 2.
        lower_final = (i == n - 1)
        first_order = self.config.solver_order == 1 or self.lower_order_nums < 1 or lower_final
        if s_t == 0.0:
            expm1_negh, ratio = -1.0, 0.0          # h = +inf
            h = math.inf
        else:
            lam_t = math.log(a_t) - math.log(s_t)
            h = lam_t - lam_s0
            expm1_negh, ratio = math.expm1(-h), s_t / s_s0
        c_x = ratio
        c_d0 = a_t * expm1_negh
        c_d1r = 0.0
        if not first_order:
            sig_s1 = float(self.sigmas[i - 1])
            a_s1, s_s1 = self._alpha_sigma(sig_s1)
            lam_s1 = math.log(a_s1) - math.log(s_s1)
            r0 = (lam_s0 - lam_s1) / h
            c_d1r = 0.5 * a_t * expm1_negh / r0
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return a_s0, s_s0, c_x, c_d0, c_d1r

    def step(self, model_output, timestep, sample, return_dict=False, **kw):
        a_s0, s_s0, c_x, c_d0, c_d1r = self.step_coefficients(timestep)
        x0 = (sample - s_s0 * model_output) / a_s0
        prev = c_x * sample - c_d0 * x0
        if c_d1r != 0.0:
            prev = prev - c_d1r * (x0 - self.model_outputs[-1])
        self.model_outputs = [self.model_outputs[-1], x0]
        return (prev,)

    def fused_cfg_step(self, eps_c, eps_u, guidance_scale: float, timestep, sample, x0_prev):
        """CUDA path: eps = w eps_c + (1-w) eps_u, then the DPM-Solver++ update, in ONE kernel
        (fd_step_cfg_dpm), in place on `sample` (fp32) and `x0_prev`.  No CPU fallback."""
        from .b200 import raw
        a_s0, s_s0, c_x, c_d0, c_d1r = self.step_coefficients(timestep)
        raw.step_cfg_dpm(eps_c, eps_u, sample, x0_prev, (guidance_scale, a_s0, s_s0, c_x, c_d0, c_d1r))
        return sample


@_host_tables
class LCMScheduler(_SchedulerBase):
    """Latent-consistency sampler: boundary-condition mix with sigma_data 0.5 and
    timestep_scaling 10, re-noising to the next timestep except on the last step."""

    def __init__(self, original_inference_steps=50, timestep_scaling=10.0, sigma_data=0.5, **kw):
        super().__init__(original_inference_steps=original_inference_steps, timestep_scaling=timestep_scaling,
                         sigma_data=sigma_data, **kw)
        self._step_index = None

    def set_timesteps(self, num_inference_steps=None, device=None, original_inference_steps=None,
                      timesteps=None, strength=1.0):
        if timesteps is not None:
            ts = np.array(timesteps, dtype=np.int64)
        else:
            orig = original_inference_steps or self.config.original_inference_steps
            k = self.config.num_train_timesteps // orig
            lcm_origin = np.asarray(list(range(1, int(orig * strength) + 1))) * k - 1
            lcm_origin = lcm_origin[::-1].copy()
            idx = np.floor(np.linspace(0, len(lcm_origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
            ts = lcm_origin[idx]
        self.num_inference_steps = len(ts)
        self.timesteps = torch.from_numpy(ts.astype(np.int64))
        self._step_index = None

    def scalings(self, timestep):
        s = float(timestep) * self.config.timestep_scaling
        sd = self.config.sigma_data
        return sd ** 2 / (s ** 2 + sd ** 2), s / (s ** 2 + sd ** 2) ** 0.5

    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        if self._step_index is None:
            self._step_index = int((self.timesteps == int(timestep)).nonzero()[0])
        i = self._step_index
        t = int(timestep)
        last = i == len(self.timesteps) - 1
        prev_t = int(self.timesteps[i + 1]) if not last else t
        ac_t = float(self.alphas_cumprod[t])
        ac_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else 1.0
        c_skip, c_out = self.scalings(t)
        x0 = (sample - (1 - ac_t) ** 0.5 * model_output) / ac_t ** 0.5
        denoised = c_out * x0 + c_skip * sample
        if not last:
            noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
            prev = ac_prev ** 0.5 * denoised + (1 - ac_prev) ** 0.5 * noise
        else:
            prev = denoised
        self._step_index += 1
        return (prev, denoised)


@_host_tables
class EulerDiscreteScheduler(_SchedulerBase):
    """diffusers `EulerDiscreteScheduler` (epsilon prediction) — the `TEACHER_SAMPLING_SCHEDULER` of the example yamls,
    used only to draw the teacher's reference samples in `log_samples` (reference flash_diffusion_model.py:866-913):
    sigma_i = sqrt((1 - abar_t) / abar_t) on the spaced timesteps, final sigma 0; the sample lives in sigma space
    (`init_noise_sigma = sqrt(sigma_max^2 + 1)` for leading/trailing spacing, `scale_model_input` divides by
    sqrt(sigma^2 + 1)); step: x += (sigma_next - sigma) * eps."""

    def __init__(self, timestep_spacing="leading", **kw):
        super().__init__(timestep_spacing=timestep_spacing, **kw)
        self._step_index = None
        self.set_timesteps(self.config.num_train_timesteps)

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            ts = np.array(timesteps, dtype=np.int64)
        else:
            ts = np.clip(_spaced_timesteps(self.config.num_train_timesteps, num_inference_steps,
                                           self.config.timestep_spacing, self.config.steps_offset),
                         0, self.config.num_train_timesteps - 1)
        self.num_inference_steps = len(ts)
        ac = self.alphas_cumprod.double().numpy()
        sig = np.sqrt((1 - ac) / ac)[ts]
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]))
        self.timesteps = torch.from_numpy(ts)
        smax = float(self.sigmas.max())
        self.init_noise_sigma = smax if self.config.timestep_spacing == "linspace" else (smax * smax + 1.0) ** 0.5
        self._step_index = None

    def _idx(self, timestep):
        if self._step_index is None:
            hits = (self.timesteps == int(timestep)).nonzero()
            self._step_index = int(hits[0]) if len(hits) else 0
        return self._step_index

    def scale_model_input(self, sample, timestep=None):
        s = float(self.sigmas[self._idx(timestep)])
        return sample / (s * s + 1.0) ** 0.5

    def step(self, model_output, timestep, sample, return_dict=False, generator=None, **kw):
        i = self._idx(timestep)
        s, s_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        prev = sample + (s_next - s) * model_output
        self._step_index = i + 1
        return (prev,)


@_host_tables
class EulerAncestralDiscreteScheduler(EulerDiscreteScheduler):
    """ancestral variant: deterministic step to sigma_down, then fresh noise of scale sigma_up."""

    def step(self, model_output, timestep, sample, return_dict=False, generator=None, **kw):
        i = self._idx(timestep)
        s, s_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        up = (s_next ** 2 * (s ** 2 - s_next ** 2) / s ** 2) ** 0.5 if s > 0 else 0.0
        down = (s_next ** 2 - up ** 2) ** 0.5
        prev = sample + (down - s) * model_output
        if up > 0:
            prev = prev + up * torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        self._step_index = i + 1
        return (prev,)


@_host_tables
class FlowMatchEulerDiscreteScheduler:
    """Rectified-flow Euler scheduler used by the SD3 recipe (reference examples/train_flash_sd3.py:123-141,
    consumed at src/flash/models/flash_sd3/flash_diffusion_model.py:253-324 and :1043-1060).

    Restated from upstream diffusers' published `FlowMatchEulerDiscreteScheduler` (v0.29): training grid
    sigma_i = shift*s/(1+(shift-1)*s), s = i/N for i = N..1, timesteps = sigma*N; `set_timesteps(K)` takes K values
    linearly spaced between sigma_max*N and sigma_min*N and applies the shift map to them AGAIN (upstream v0.29
    behaviour, kept); `step` is x + (sigma_next - sigma) * v with a trailing sigma of 0.

    `timestep_spacing="trailing"` exists only in the authors' diffusers fork, which is not available: it is read here
    as the trailing grid of the other schedulers (raw timesteps N, N-N/K, ..., N/K, shift map applied once).  That
    reading is an ASSUMPTION and is stated in DESIGN.md; the default spacing follows upstream.
    """
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0, timestep_spacing="linspace", **extra):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                      timestep_spacing=timestep_spacing, **extra)
        n = num_train_timesteps
        sig = torch.from_numpy(np.linspace(1, n, n, dtype=np.float32)[::-1].copy()) / n
        sig = self._shift(sig)
        self.timesteps = sig * n
        self.sigmas = sig.clone()
        self.sigma_min, self.sigma_max = float(sig[-1]), float(sig[0])
        self.num_inference_steps = None
        self._step_index = None

    def _shift(self, s):
        k = self.config.shift
        return k * s / (1 + (k - 1) * s)

    @classmethod
    def from_pretrained(cls, repo: str = None, subfolder: str = None, revision: str = None, **overrides):
        cfg = dict(num_train_timesteps=1000, shift=3.0)          # stabilityai/stable-diffusion-3-medium scheduler
        cfg.update(overrides)
        return cls(**cfg)

    @classmethod
    def from_config(cls, config, **overrides):
        cfg = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        cfg.update(overrides)
        return cls(**cfg)

    def set_timesteps(self, num_inference_steps=None, device=None):
        n = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        if self.config.timestep_spacing == "trailing":
            raw = np.arange(n, 0, -n / num_inference_steps, dtype=np.float64)[:num_inference_steps]
            sig = self._shift(torch.from_numpy((raw / n).astype(np.float32)))
        else:
            raw = np.linspace(self.sigma_max * n, self.sigma_min * n, num_inference_steps, dtype=np.float32)
            sig = self._shift(torch.from_numpy(raw / n))
        self.timesteps = (sig * n).to(device) if device is not None else sig * n
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self._step_index = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def index_for(self, timestep) -> int:
        hits = (self.timesteps == float(timestep)).nonzero()
        if hits.numel() == 0:
            raise ValueError(f"timestep {float(timestep)} is not on the current schedule")
        return int(hits[0])

    def scale_noise(self, sample, timestep, noise):
        s = float(self.sigmas[self.index_for(timestep)])
        return s * noise + (1.0 - s) * sample

    def step(self, model_output, timestep, sample, return_dict=False, **kw):
        i = self.index_for(timestep)
        s, s_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        prev = sample + (s_next - s) * model_output
        self._step_index = i + 1
        return (prev,)

    def fused_cfg_step(self, v_c, v_u, guidance_scale: float, timestep, sample, scratch):
        """CUDA path: v = w v_c + (1-w) v_u and the Euler update in ONE kernel (fd_step_cfg_dpm), in place on `sample`
        (fp32).  The kernel computes x0 = (x - s v)/a and x' = c_x x - c_d0 x0 - c_d1r (x0 - x0_prev); with a = 1,
        s = sigma it is the flow's x0 read-out, and x + (sigma' - sigma) v = (sigma'/sigma) x + (1 - sigma'/sigma) x0."""
        from .b200 import raw
        i = self.index_for(timestep)
        s, s_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        r = s_next / s
        raw.step_cfg_dpm(v_c, v_u, sample, scratch, (guidance_scale, 1.0, s, r, -(1.0 - r), 0.0))
        self._step_index = i + 1
        return sample


@_host_tables
class FlashFlowMatchEulerDiscreteScheduler(FlowMatchEulerDiscreteScheduler):
    """Few-step student sampler of the SD3 recipe (reference examples/configs/flash_sd3.yaml `SAMPLING_SCHEDULER`,
    consumed at flash_sd3/flash_diffusion_model.py:694-790).  The class exists only in the authors' diffusers fork;
    it is read here as the flow-matching analogue of the LCM sampler the other recipes use: the student's velocity
    gives x0 = x - sigma*v (exactly the `student_output` the objective trains, :324), which is re-noised to the next
    level, x' = (1 - sigma')*x0 + sigma'*noise.  ASSUMPTION, stated in DESIGN.md."""

    def step(self, model_output, timestep, sample, generator=None, return_dict=False, **kw):
        i = self.index_for(timestep)
        s, s_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        x0 = sample - s * model_output
        if s_next > 0:
            noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
            prev = (1.0 - s_next) * x0 + s_next * noise
        else:
            prev = x0
        self._step_index = i + 1
        return (prev, x0)
