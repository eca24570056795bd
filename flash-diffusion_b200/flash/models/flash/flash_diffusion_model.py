"""Flash-Diffusion distillation objective — B200 host side.

Mirrors reference `FlashDiffusion` (src/flash/models/flash/flash_diffusion_model.py:38-1019): same
constructor, `forward(batch, batch_idx, step) -> {"loss": [loss_G_total, loss_D], "teacher_output",
"student_output", "noisy_sample", "start_timestep"}`, `sample()`, and the same RNG draw order
(SURVEY.md §8d).  The denoisers behind it are the hand-written-kernel wrappers of flash.models.unets.

Output-preserving restructurings for B200 (each cited to the reference lines it replaces):
  * teacher cond / uncond evaluated as ONE call at batch 2B instead of two calls (:297-313, :432-444);
  * CFG combine + DPM-Solver++ update in one fused kernel (`fd_step_cfg_dpm`) (:316-324);
  * the three conditioner passes are deduplicated when no conditioner has a UCG rate (:188-205);
  * every random draw can be injected through `draws=` so that oracle and kernels see identical values
    (SURVEY.md §8c decision 5); without it the global generator is consumed in the reference's order.
"""
import logging
from copy import deepcopy
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from ..base.base_model import BaseModel
from ..utils import append_dims, extract_into_tensor
from .flash_diffusion_config import FlashDiffusionConfig


def gaussian_mixture(k, locs, var, mode_probs=None):
    """pmf kernel of the start-index mixture (reference :23-35)."""
    if mode_probs is None:
        mode_probs = [1 / len(locs)] * len(locs)

    def _pdf(x):
        return sum(mode_probs[i] * torch.exp(-torch.tensor([(x - loc) ** 2 / var])) for i, loc in enumerate(locs))

    return _pdf


def _cat_conditioning(a, b):
    return {"cond": {k: torch.cat([a["cond"][k], b["cond"][k]], dim=0) for k in a["cond"]}}


class FlashDiffusion(BaseModel):
    def __init__(self, config: FlashDiffusionConfig, student_denoiser, teacher_denoiser=None,
                 teacher_noise_scheduler=None, teacher_sampling_noise_scheduler=None,
                 sampling_noise_scheduler=None, vae=None, conditioner=None, adapter=None,
                 discriminator: torch.nn.Module = None):
        super().__init__(config)
        self.student_denoiser = student_denoiser
        self.teacher_denoiser = teacher_denoiser
        self.teacher_noise_scheduler = teacher_noise_scheduler
        self.teacher_sampling_noise_scheduler = teacher_sampling_noise_scheduler
        self.sampling_noise_scheduler = sampling_noise_scheduler
        self.vae = vae
        self.conditioner = conditioner
        self.adapter = adapter
        for name in ("guidance_scale_min", "guidance_scale_max", "ucg_keys", "adapter_input_key", "K",
                     "num_iterations_per_K", "distill_loss_type", "timestep_distribution",
                     "mixture_num_components", "mixture_var", "adapter_conditioning_scale", "use_dmd_loss",
                     "dmd_loss_scale", "distill_loss_scale", "adversarial_loss_scale", "gan_loss_type",
                     "mode_probs", "use_teacher_as_real", "use_empty_prompt"):
            setattr(self, name, getattr(config, name))
        self.iter_steps = 0
        self.discriminator = discriminator
        self.disc_update_counter = 0
        self.switch_teacher = False          # reference reads an attribute it never sets
        if self.discriminator is None:
            logging.warning("No discriminator provided. Adversarial loss will be ignored.")
            self.use_adversarial_loss = False
        self.disc_backbone = self.teacher_denoiser
        if self.distill_loss_type == "lpips":
            # reference :102-103 `lpips.LPIPS(net="vgg")`; here the VGG16 stack runs on the B200 conv kernels.  The
            # published VGG / lin weights are not available offline: random-init unless a state dict is loaded.
            from ..lpips import LPIPS
            if self.vae is None:
                raise ValueError("distill_loss_type='lpips' decodes the latents: a VAE is required (reference :393-394)")
            self.lpips = LPIPS(net="vgg")
        if adapter is not None:
            raise NotImplementedError("T2I adapters are out of scope of the B200 hot path (SURVEY.md §2 row 7)")
        self.K_steps = np.cumsum(self.num_iterations_per_K)
        self.K_prev = self.K[0]
        if teacher_noise_scheduler is not None:
            if hasattr(teacher_noise_scheduler, "alphas_cumprod"):
                ac = teacher_noise_scheduler.alphas_cumprod
                self.register_buffer("sqrt_alpha_cumprod", torch.sqrt(ac))
                self.register_buffer("sigmas", torch.sqrt(1 - ac))
            elif hasattr(teacher_noise_scheduler, "sigmas"):
                self.register_buffer("sqrt_alpha_cumprod", torch.sqrt(1 - teacher_noise_scheduler.sigmas ** 2))
                self.register_buffer("sigmas", teacher_noise_scheduler.sigmas)
        # B200 options
        self.use_cuda_graphs = True      # replay frozen-teacher evaluations from a CUDA graph (same kernels)
        self.__dict__["_graphed"] = {}
        # Output-preserving dead-work elision: on the discriminator turn (odd `step`) the
        # trainer back-propagates ONLY loss[1] (reference trainer.py:213-217), which depends on the detached student
        # output, the GAN backbone features and the discriminator — not on the teacher rollout, the distillation loss
        # or the DMD loss the reference recomputes and discards there.  Off by default (strict reference schedule);
        # when on, those are skipped (loss[0] is returned as None) and the student runs without an autograd graph.
        self.elide_unused_generator_pass = False
        self.batch_cfg = True            # cond+uncond as one 2B call (output-preserving)
        self.cache_teacher_kv = True     # cross-attention K/V of the text conditioning computed once per rollout
        self.dedupe_conditioning = True  # one conditioner pass when every ucg_rate is 0 (output-preserving)

    # ------------------------------------------------------------------ helpers (reference :127-177,:687-752)
    def _encode_inputs(self, batch):
        with torch.no_grad():
            return self.vae.encode(batch[self.vae.config.input_key])

    def _start_index_pmf(self, K: int, K_step: int) -> torch.Tensor:
        if self.timestep_distribution == "uniform":
            return torch.ones(K) / K
        if self.timestep_distribution == "gaussian":
            prob = torch.tensor([float(torch.exp(-torch.tensor([(i - K / 2) ** 2 / K]))) for i in range(K)])
            return prob / prob.sum()
        m = self.mixture_num_components[K_step]
        locs = [i * (K // m) for i in range(m)]
        pdf = gaussian_mixture(K, locs=locs, var=self.mixture_var[K_step], mode_probs=self.mode_probs[K_step])
        prob = torch.tensor([float(pdf(i)) for i in range(K)])
        return prob / prob.sum()

    def _get_timesteps(self, num_samples=1, K=1, K_step=1, device="cpu", start_idx=None):
        self.teacher_noise_scheduler.set_timesteps(K)
        if start_idx is None:
            start_idx = torch.multinomial(self._start_index_pmf(K, K_step), 1)
        else:
            start_idx = torch.as_tensor([int(start_idx)])
        start_timestep = self.teacher_noise_scheduler.timesteps[start_idx].to(device).repeat(num_samples)
        return start_idx, start_timestep

    def _get_conditioning(self, batch, ucg_keys: List[str] = None, set_ucg_rate_zero=False, *args, **kwargs):
        if self.conditioner is None:
            return None
        return self.conditioner(batch, ucg_keys=ucg_keys, set_ucg_rate_zero=set_ucg_rate_zero, vae=self.vae,
                                *args, **kwargs)

    def _scalings_for_boundary_conditions(self, timestep, sigma_data=0.5):
        s = timestep / 0.1
        return sigma_data ** 2 / (s ** 2 + sigma_data ** 2), s / (s ** 2 + sigma_data ** 2) ** 0.5

    def _predicted_x_0(self, model_output, timesteps, sample, prediction_type, alphas, sigmas, input_sample):
        if prediction_type == "epsilon":
            sig = extract_into_tensor(sigmas, timesteps, sample.shape)
            alp = extract_into_tensor(alphas, timesteps, sample.shape)
            safe = torch.where(alp > 0, alp, torch.ones_like(alp))
            x0 = (sample - sig * model_output) / safe
            return torch.where(alp > 0, x0, torch.where(alp == 0, input_sample, torch.zeros_like(x0)))
        if prediction_type == "v_prediction":
            sig = extract_into_tensor(sigmas, timesteps, sample.shape)
            alp = extract_into_tensor(alphas, timesteps, sample.shape)
            return alp * sample - sig * model_output
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")

    def _ucg_is_deterministic(self):
        return self.conditioner is None or all(c.ucg_rate == 0 for c in self.conditioner.conditioners)

    def _call_frozen(self, denoiser, sample, timestep, conditioning, clone=True, **kw):
        """Frozen-denoiser evaluation; on CUDA it is replayed from a CUDA graph (flash.b200.graphs)."""
        if self.use_cuda_graphs and sample.is_cuda:
            from ...b200.graphs import GraphedDenoiser
            if GraphedDenoiser.eligible(denoiser, sample):
                g = self.__dict__["_graphed"].get(id(denoiser))
                if g is None:
                    g = self.__dict__["_graphed"][id(denoiser)] = GraphedDenoiser(denoiser)
                return g(sample, timestep, conditioning, clone=clone, **kw)
        return denoiser(sample=sample, timestep=timestep, conditioning=conditioning, **kw)

    def _teacher_pair(self, denoiser, sample, timestep, cond, uncond, clone=True, **kw):
        """eps_cond, eps_uncond of the frozen teacher; one 2B call when batching is on.  Consecutive calls with the same
        conditioning objects (the K-step rollout, then the DMD teacher pair) re-use the cross-attention K/V projections
        of the first one (`kv_cache`, output-preserving: they depend on the text conditioning only)."""
        if self.batch_cfg and cond is not None:
            B = sample.shape[0]
            if (getattr(self, "cache_teacher_kv", False) and getattr(denoiser, "supports_kv_cache", False) and sample.is_cuda
                    and not torch.is_grad_enabled() and "crossattn" in cond["cond"]):
                key = (id(denoiser), id(cond["cond"]["crossattn"]), id(uncond["cond"]["crossattn"]), B)
                kw = dict(kw, kv_cache="reuse" if self.__dict__.get("_kv_key") == key else "fill")
                self.__dict__["_kv_key"] = key
            both = self._call_frozen(denoiser, torch.cat([sample, sample], dim=0),
                                     torch.cat([timestep, timestep], dim=0), _cat_conditioning(cond, uncond),
                                     clone=clone, **kw)
            return both[:B], both[B:]
        return (denoiser(sample=sample, timestep=timestep, conditioning=cond, **kw),
                denoiser(sample=sample, timestep=timestep, conditioning=uncond, **kw))

    # ------------------------------------------------------------------ forward (reference :179-366)
    def forward(self, batch: Dict[str, Any], batch_idx=0, step=0, draws: Optional[Dict[str, Any]] = None,
                *args, **kwargs):
        draws = draws or {}
        device = kwargs.pop("device", None)
        ckw = {} if device is None else {"device": device}     # the text conditioners take the device (reference :188-205)
        self.iter_steps += 1
        self.__dict__["_kv_key"] = None          # new batch: cached teacher K/V are stale
        z = self._encode_inputs(batch) if self.vae is not None else batch[self.input_key]

        conditioning = self._get_conditioning(batch, set_ucg_rate_zero=True, **ckw)
        if self.dedupe_conditioning and self._ucg_is_deterministic():
            student_conditioning = conditioning
        else:
            student_conditioning = self._get_conditioning(batch, **ckw)
        if self.use_empty_prompt and "text" in self.ucg_keys:
            uncond_batch = dict(batch)
            uncond_batch["text"] = [""] * len(batch["text"])
            unconditional_conditioning = self._get_conditioning(uncond_batch, set_ucg_rate_zero=True, **ckw)
        else:
            unconditional_conditioning = self._get_conditioning(batch, ucg_keys=self.ucg_keys, **ckw)

        if self.iter_steps > self.K_steps[-1]:
            K_step = len(self.K) - 1
        else:
            K_step = int(np.argmax(self.iter_steps < self.K_steps))
        K = self.K[K_step]
        g_min, g_max = self.guidance_scale_min[K_step], self.guidance_scale_max[K_step]
        if K != self.K_prev:
            self.K_prev = K
            if self.switch_teacher:
                self.teacher_denoiser = deepcopy(self.student_denoiser)
                self.teacher_denoiser.freeze()

        sched = self.teacher_noise_scheduler
        noise = draws["noise"] if "noise" in draws else torch.randn_like(z)
        start_idx, start_timestep = self._get_timesteps(z.shape[0], K=K, K_step=K_step, device=z.device,
                                                        start_idx=draws.get("start_idx"))
        if int(start_idx) == 0:
            noisy_sample_init = noise * sched.init_noise_sigma
        else:
            noisy_sample_init = sched.add_noise(z, noise, start_timestep)
        student_in = sched.scale_model_input(noisy_sample_init, start_timestep)
        start_t_host = int(sched.timesteps[int(start_idx)])          # host copy: no device sync for the return value

        # guidance scale: drawn on the host exactly like the reference (`torch.rand(1).to(device)`, :284-286) but kept
        # as a host float so that the fused rollout never has to read it back from the device
        if "guidance" in draws:
            guidance_host = float(draws["guidance"])
        else:
            guidance_host = float(torch.rand(1)) * (g_max - g_min) + g_min

        lean = (self.elide_unused_generator_pass and step % 2 == 1 and not self.use_teacher_as_real
                and self.discriminator is not None)
        # The frozen-teacher rollout (CUDA-graph replays: GPU-heavy, host-light) is ENQUEUED BEFORE the student
        # forward (eager launches + autograd bookkeeping: host-heavy): the two are independent (:260-265 vs :288-324),
        # so the host prepares the student's launches while the GPU is busy with the rollout.  Same values, same
        # random-draw order.
        teacher_output = None
        if not lean:
            teacher_output = self._teacher_rollout(noisy_sample_init, conditioning, unconditional_conditioning,
                                                   int(start_idx), guidance_host)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not lean):
            student_noise_pred = self.student_denoiser(sample=student_in, timestep=start_timestep,
                                                       conditioning=student_conditioning)
        c_skip, c_out = self._scalings_for_boundary_conditions(start_timestep)
        c_skip, c_out = append_dims(c_skip, z.ndim), append_dims(c_out, z.ndim)
        student_x0 = self._predicted_x_0(student_noise_pred, start_timestep.type(torch.int64), noisy_sample_init,
                                         "epsilon", self.sqrt_alpha_cumprod, self.sigmas, z)

        student_output = c_skip * noisy_sample_init + c_out * student_x0
        if lean:
            gan_loss = self._gan_loss(z, batch, student_output, None, conditioning, None, step=step, draws=draws)
            return {"loss": [None, gan_loss[1]], "teacher_output": None, "student_output": student_output,
                    "noisy_sample": noisy_sample_init, "start_timestep": start_t_host}

        distill = self._distill_loss(student_output, teacher_output)
        loss = distill * self.distill_loss_scale[K_step]
        dmd = None
        if self.use_dmd_loss:
            dmd = self._dmd_loss(student_output, student_conditioning, conditioning, unconditional_conditioning, None,
                                 K, K_step, draws)
            loss = loss + dmd * self.dmd_loss_scale[K_step]
        gan_loss = self._gan_loss(z, batch, student_output, teacher_output, conditioning, None, step=step, draws=draws)
        loss = loss + self.adversarial_loss_scale[K_step] * gan_loss[0]
        # un-scaled terms of the objective (detached; parity tests compare each against the oracle)
        self.__dict__["last_loss_terms"] = {
            k: (v.detach() if torch.is_tensor(v) else v)
            for k, v in dict(distill=distill, dmd=dmd, gan_G=gan_loss[0], gan_D=gan_loss[1]).items()}
        return {"loss": [loss, gan_loss[1]], "teacher_output": teacher_output, "student_output": student_output,
                "noisy_sample": noisy_sample_init, "start_timestep": start_t_host}

    @torch.no_grad()
    def _teacher_rollout(self, noisy_sample_init, conditioning, unconditional_conditioning, start_idx, guidance_scale):
        sched = self.teacher_noise_scheduler
        x = noisy_sample_init.clone().detach()
        B = x.shape[0]
        fused = x.is_cuda and hasattr(sched, "fused_cfg_step")
        w = float(guidance_scale)
        if fused:
            x = x.float().contiguous()
            x0_prev = torch.zeros_like(x)
        for t in sched.timesteps[start_idx:]:
            timestep = torch.tensor([t], device=x.device).repeat(B)
            x_in = sched.scale_model_input(x, t)
            eps_c, eps_u = self._teacher_pair(self.teacher_denoiser, x_in, timestep, conditioning,
                                              unconditional_conditioning, clone=not fused)
            if fused:
                sched.fused_cfg_step(eps_c.contiguous(), eps_u.contiguous(), w, t, x, x0_prev)
            else:
                eps = w * eps_c + (1 - w) * eps_u
                x = sched.step(eps, t, x, return_dict=False)[0]
        return x

    # ------------------------------------------------------------------ losses (reference :368-667)
    def _distill_loss(self, student_output, teacher_output):
        if self.distill_loss_type == "l2":
            return torch.mean(((student_output - teacher_output) ** 2).reshape(student_output.shape[0], -1), 1).mean()
        if self.distill_loss_type == "l1":
            return torch.mean(torch.abs(student_output - teacher_output).reshape(student_output.shape[0], -1), 1).mean()
        if self.distill_loss_type == "lpips":
            # reference :383-397 — center crop 64x64 latents, decode both, clamp, LPIPS-VGG, mean
            ch = (student_output.shape[2] - 64) // 2
            cw = (student_output.shape[3] - 64) // 2
            s_crop = student_output[:, :, ch:ch + 64, cw:cw + 64]
            t_crop = teacher_output[:, :, ch:ch + 64, cw:cw + 64]
            decoded_student = self.vae.decode(s_crop).clamp(-1, 1)
            with torch.no_grad():
                decoded_teacher = self.vae.decode(t_crop).clamp(-1, 1)
            return self.lpips(decoded_student, decoded_teacher).mean()
        raise NotImplementedError(f"Loss type {self.distill_loss_type} not implemented")

    def _dmd_loss(self, student_output, student_conditioning, conditioning, unconditional_conditioning,
                  down_intrablock_additional_residuals, K, K_step, draws=None):
        draws = draws or {}
        sched = self.teacher_noise_scheduler
        noise = draws["dmd_noise"] if "dmd_noise" in draws else torch.randn_like(student_output)
        if "dmd_timestep" in draws:
            timestep = draws["dmd_timestep"].to(student_output.device)
        else:
            timestep = torch.randint(0, sched.config.num_train_timesteps, (student_output.shape[0],),
                                     device=student_output.device)
        noisy_student = sched.add_noise(student_output, noise, timestep)
        with torch.no_grad():
            real_c, real_u = self._teacher_pair(self.teacher_denoiser, noisy_student, timestep, conditioning,
                                                unconditional_conditioning)
            fake = self.student_denoiser(sample=noisy_student, timestep=timestep, conditioning=student_conditioning)
            if "dmd_guidance" in draws:
                w = torch.as_tensor([float(draws["dmd_guidance"])], device=student_output.device)
            else:
                w = (torch.rand(1).to(student_output.device)
                     * (self.guidance_scale_max[K_step] - self.guidance_scale_min[K_step]) + self.guidance_scale_min[K_step])
        real = w * real_c + (1 - w) * real_u
        alpha_prod_t = sched.alphas_cumprod.to(device=student_output.device, dtype=student_output.dtype)[timestep]
        beta_prod_t = 1.0 - alpha_prod_t
        coeff = ((-fake) - (-real)) * beta_prod_t.view(-1, 1, 1, 1) ** 0.5 / alpha_prod_t.view(-1, 1, 1, 1) ** 0.5
        pred_x0 = self._predicted_x_0(real, timestep, noisy_student, "epsilon", self.sqrt_alpha_cumprod, self.sigmas,
                                      student_output)
        weight = 1.0 / ((student_output - pred_x0).abs().mean([1, 2, 3], keepdim=True) + 1e-5).detach()
        return F.mse_loss(student_output, (student_output - weight * coeff).detach(), reduction="mean")

    def _gan_loss(self, z, batch, student_output, teacher_output, conditioning,
                  down_intrablock_additional_residuals=None, step=0, draws=None):
        draws = draws or {}
        self.disc_update_counter += 1
        sched = self.teacher_noise_scheduler
        noise = draws["gan_noise"] if "gan_noise" in draws else torch.randn_like(student_output)
        real = teacher_output if self.use_teacher_as_real else z
        B = student_output.shape[0]
        if "gan_timesteps" in draws:
            timesteps = draws["gan_timesteps"].to(student_output.device)
        else:
            idx = torch.tensor([0.25] * 4).multinomial(B, replacement=True).to(student_output.device)
            timesteps = torch.tensor([10, 250, 500, 750], device=student_output.device, dtype=torch.long)[idx]
        generator_turn = step % 2 == 0
        fake_in = student_output if generator_turn else student_output.detach()
        noisy_fake = sched.add_noise(fake_in, noise, timesteps)
        noisy_real = sched.add_noise(real, noise, timesteps)
        noisy_sample = torch.cat([noisy_fake, noisy_real], dim=0)
        cond2 = None
        if conditioning is not None:
            cond2 = {"cond": {k: torch.cat([v, v], dim=0) for k, v in conditioning["cond"].items()}}
        # frozen teacher backbone -> mid-block features (reference :563-569); on the discriminator turn the
        # reference detaches the fake features (:600), so no graph is needed there at all
        with torch.set_grad_enabled(generator_turn and torch.is_grad_enabled()):
            feats = self._call_frozen(self.disc_backbone, noisy_sample, torch.cat([timesteps, timesteps], dim=0),
                                      cond2, return_intermediate=True)
        f_fake, f_real = feats.chunk(2, dim=0)
        return self._gan_objective(f_fake, f_real, B, generator_turn)

    def _gan_objective(self, f_fake, f_real, B, generator_turn):
        """[loss_G, loss_D] of the configured GAN type on backbone features (reference :571-667; the SD3 variant
        flash_sd3/flash_diffusion_model.py:571-658 is the same table)."""
        D = self.discriminator
        dev = f_fake.device
        valid = torch.ones(B, 1, device=dev)
        fake_t = torch.zeros(B, 1, device=dev)
        t = self.gan_loss_type
        if t == "wgan":
            for p in D.parameters():
                p.data.clamp_(-0.01, 0.01)
        if generator_turn:
            d_f = D(f_fake)
            if t in ("wgan", "hinge"):
                loss_G = -d_f.mean()
            elif t == "lsgan":
                loss_G = F.mse_loss(torch.sigmoid(d_f), valid)
            elif t == "non-saturating":
                loss_G = -torch.mean(torch.log(torch.sigmoid(d_f) + 1e-8))
            else:
                loss_G = F.binary_cross_entropy_with_logits(d_f, valid)
            return [loss_G, 0]
        d_r, d_f = D(f_real.detach()), D(f_fake.detach())
        if t == "wgan":
            loss_D = -d_r.mean() + d_f.mean()
        elif t == "lsgan":
            loss_D = 0.5 * (F.mse_loss(torch.sigmoid(d_r), valid) + F.mse_loss(torch.sigmoid(d_f), fake_t))
        elif t == "hinge":
            loss_D = F.relu(1.0 - d_r).mean() + F.relu(1.0 + d_f).mean()
        elif t == "non-saturating":
            loss_D = -torch.mean(torch.log(torch.sigmoid(d_r) + 1e-8) + torch.log(1 - torch.sigmoid(d_f) + 1e-8))
        else:
            loss_D = F.binary_cross_entropy_with_logits(d_r, valid) + F.binary_cross_entropy_with_logits(d_f, fake_t)
        return [0, loss_D]

    # ------------------------------------------------------------------ few-step sampler (reference :754-915)
    @torch.no_grad()
    def sample(self, z, num_steps=20, guidance_scale=1.0, teacher_guidance_scale=5.0, conditioner_inputs=None,
               uncond_conditioner_inputs=None, max_samples=None, verbose=False, log_teacher_samples=False,
               adapter_conditioning_scale=1.0, generator=None):
        self.__dict__["_kv_key"] = None          # cached K/V never outlive one call
        self.teacher_noise_scheduler.set_timesteps(num_steps)
        try:
            self.sampling_noise_scheduler.set_timesteps(timesteps=self.teacher_noise_scheduler.timesteps)
        except Exception:
            self.sampling_noise_scheduler.set_timesteps(num_steps)
        sample = z
        conditioning = self._get_conditioning(conditioner_inputs, set_ucg_rate_zero=True, device=z.device)
        if uncond_conditioner_inputs is not None:
            unconditional = self._get_conditioning(uncond_conditioner_inputs, set_ucg_rate_zero=True, device=z.device)
        else:
            unconditional = self._get_conditioning(conditioner_inputs, ucg_keys=self.ucg_keys, device=z.device)
        if max_samples is not None:
            sample = sample[:max_samples]
            if conditioning:
                conditioning["cond"] = {k: v[:max_samples] for k, v in conditioning["cond"].items()}
                unconditional["cond"] = {k: v[:max_samples] for k, v in unconditional["cond"].items()}
        sample_init = sample
        sched = self.sampling_noise_scheduler
        sample = sample * sched.init_noise_sigma
        for t in sched.timesteps:
            x_in = sched.scale_model_input(sample, t)
            ts = t.to(z.device).repeat(x_in.shape[0])
            eps_c, eps_u = self._teacher_pair(self.student_denoiser, x_in, ts, conditioning, unconditional)
            eps = guidance_scale * eps_c + (1 - guidance_scale) * eps_u
            kw = {"generator": generator} if generator is not None else {}
            sample = sched.step(eps, t, sample, return_dict=False, **kw)[0]
        decoded = self.vae.decode(sample) if self.vae is not None else sample
        decoded_ref = None
        if log_teacher_samples:
            ts_sched = self.teacher_sampling_noise_scheduler
            ts_sched.set_timesteps(num_steps)
            ref = sample_init * ts_sched.init_noise_sigma
            for t in ts_sched.timesteps:
                x_in = ts_sched.scale_model_input(ref, t)
                ts = t.to(z.device).repeat(x_in.shape[0])
                eps_c, eps_u = self._teacher_pair(self.teacher_denoiser, x_in, ts, conditioning, unconditional)
                eps = teacher_guidance_scale * eps_c + (1 - teacher_guidance_scale) * eps_u
                ref = ts_sched.step(eps, t, ref, return_dict=False)[0]
            decoded_ref = self.vae.decode(ref) if self.vae is not None else ref
        self.__dict__["_kv_key"] = None
        return decoded, decoded_ref

    # ------------------------------------------------------------------ sample logging (reference :917-1019)
    def log_samples(self, batch: Dict[str, Any], input_shape=None, guidance_scale: float = 1.0,
                    teacher_guidance_scale: float = 5.0, max_samples: int = 8, num_steps=20, device="cpu",
                    log_teacher_samples=False, conditioner_inputs: Dict = None, conditioner_uncond_inputs: Dict = None,
                    adapter_conditioning_scale: float = 1.0, **sample_kwargs):
        """{"samples_{n}_steps/{SamplerClass}_{cfg}_cfg/student": tensor, ".../teacher": tensor} for every n in
        `num_steps`; the number of samples is capped by `max_samples` and by the shortest conditioning entry."""
        return self._log_samples(batch, input_shape, guidance_scale, teacher_guidance_scale, max_samples, num_steps,
                                 device, log_teacher_samples, conditioner_inputs, conditioner_uncond_inputs,
                                 adapter_conditioning_scale=adapter_conditioning_scale, **sample_kwargs)

    def _log_samples(self, batch, input_shape, guidance_scale, teacher_guidance_scale, max_samples, num_steps, device,
                     log_teacher_samples, conditioner_inputs, conditioner_uncond_inputs, **sample_kwargs):
        steps = [num_steps] if isinstance(num_steps, int) else list(num_steps)
        logs = {}
        N = max_samples
        if batch is not None:
            N = min(N, min(len(batch[key]) for key in batch))

        def merge(extra):
            nonlocal N
            N = min(N, min(len(extra[key]) for key in extra))
            extra.update({k: v.to(device) for k, v in extra.items() if isinstance(v, torch.Tensor)})
            return extra

        if conditioner_inputs is not None:
            batch.update(merge(conditioner_inputs))
        batch_uncond = None
        if conditioner_uncond_inputs is not None:
            batch_uncond = deepcopy(batch)
            batch_uncond.update(merge(conditioner_uncond_inputs))
        if input_shape is None:
            if self.vae is None:
                raise ValueError("input_shape must be passed when no VAE is used in the model")
            px = batch[self.vae.config.input_key].shape[2:]
            f = self.vae.downsampling_factor
            input_shape = (self.vae.latent_channels, px[0] // f, px[1] // f)
        for n in steps:
            z = torch.randn(N, *input_shape).to(device)
            logging.debug(f"Sampling {N} samples: steps={n}, guidance_scale={guidance_scale}")
            samples, samples_ref = self.sample(z, num_steps=n, conditioner_inputs=batch,
                                               uncond_conditioner_inputs=batch_uncond, guidance_scale=guidance_scale,
                                               teacher_guidance_scale=teacher_guidance_scale, max_samples=N,
                                               log_teacher_samples=log_teacher_samples, **sample_kwargs)
            logs[f"samples_{n}_steps/{self.sampling_noise_scheduler.__class__.__name__}_{guidance_scale}_cfg/student"] = samples
            if samples_ref is not None:
                logs[f"samples_{n}_steps/{self.teacher_sampling_noise_scheduler.__class__.__name__}"
                     f"_{teacher_guidance_scale}_cfg/teacher"] = samples_ref
        return logs
