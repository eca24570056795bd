"""Flash-Diffusion objective for rectified-flow (SD3) denoisers — B200 host side.

Mirrors reference `FlashDiffusionSD3` (src/flash/models/flash_sd3/flash_diffusion_model.py:38-1061): same constructor,
`forward(batch, batch_idx, step) -> {"loss": [loss_G_total, loss_D] | loss, "teacher_output", "student_output",
"noisy_sample", "start_timestep"}`, `sample()`, `get_sigmas()`, and the same order of random draws.  What changes
against the epsilon-prediction objective (flash/flash_diffusion_model.py):
  * noising is the flow interpolation  x_t = sigma*noise + (1-sigma)*z            (:258-266)
  * the student output is the velocity's x0 read-out  x_t - sigma*v               (:324)
  * the teacher rollout is Euler on the flow ODE                                   (:288-320)
  * DMD uses the unscaled score difference and the guided teacher velocity as the "x0" of its weight (:476-494)
  * DMD / GAN timesteps are looked up on the 1000-point training grid of a scheduler COPY made at construction
    (:106, :435-441, :523-528); the GAN "features" are the full backbone output because the MMDiT wrapper swallows
    `return_post_mid_blocks` (:560-565, transformers/tranformers.py:113-150)

Output-preserving restructurings for B200, as in `FlashDiffusion`: cond/uncond teacher evaluations as ONE call at
batch 2B; CFG combine + Euler update in the fused `fd_step_cfg_dpm` kernel on CUDA; the frozen teacher replayed from
a CUDA graph; sigma look-ups done on the host copies of the schedule (no device round trip); every random draw
injectable through `draws=` (tests feed oracle and product the same values).

Prompt embeddings: the reference calls `pipeline.encode_prompt` (three text encoders of the diffusers SD3 pipeline,
:199-219).  Text encoders are outside the B200 hot path, so `pipeline` is duck-typed (anything with `encode_prompt`
and `to`), and when it is None the four embedding tensors are read from the batch under the names the pipeline
returns (`prompt_embeds`, `negative_prompt_embeds`, `pooled_prompt_embeds`, `negative_pooled_prompt_embeds`).

On CUDA the denoisers are the B200 MMDiT wrappers (flash.models.transformers.sd3): teacher rollout, DMD / GAN
evaluations, the student forward and its LoRA backward all run on the hand-written kernels; on CPU (tests/) the class is
exercised around the fp32 oracle MMDiT.
"""
import logging
from copy import deepcopy
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from ..flash.flash_diffusion_model import FlashDiffusion
from .flash_diffusion_config import FlashDiffusionSD3Config

# the fixed negative prompt of the reference objective (:207-209)
NEGATIVE_PROMPT = ("deformed, distorted, disfigured, poorly drawn, bad anatomy, wrong anatomy, extra limb, missing limb, "
                   "floating limbs, mutated hands and fingers, disconnected limbs, mutation, mutated, ugly, disgusting, "
                   "blurry, amputation, NSFW")
_EMBED_KEYS = ("prompt_embeds", "negative_prompt_embeds", "pooled_prompt_embeds", "negative_pooled_prompt_embeds")


class FlashDiffusionSD3(FlashDiffusion):
    def __init__(self, config: FlashDiffusionSD3Config, student_denoiser, teacher_denoiser=None,
                 teacher_noise_scheduler=None, teacher_sampling_noise_scheduler=None, sampling_noise_scheduler=None,
                 vae=None, conditioner=None, discriminator: torch.nn.Module = None, pipeline=None,
                 cpu_offload: bool = False):
        torch.nn.Module.__init__(self)
        self.config = config
        self.input_key = config.input_key
        self.student_denoiser = student_denoiser
        self.teacher_denoiser = teacher_denoiser
        self.teacher_noise_scheduler = teacher_noise_scheduler
        self.teacher_sampling_noise_scheduler = teacher_sampling_noise_scheduler
        self.sampling_noise_scheduler = sampling_noise_scheduler
        self.vae = vae
        self.conditioner = conditioner
        for name in ("guidance_scale_min", "guidance_scale_max", "K", "num_iterations_per_K", "distill_loss_type",
                     "timestep_distribution", "mixture_num_components", "mixture_var", "use_dmd_loss",
                     "dmd_loss_scale", "distill_loss_scale", "adversarial_loss_scale", "gan_loss_type", "mode_probs",
                     "use_teacher_as_real"):
            setattr(self, name, getattr(config, name))
        self.iter_steps = 0
        self.discriminator = discriminator
        self.pipeline = pipeline
        self.cpu_offload = cpu_offload
        self.teacher_noise_scheduler_copy = deepcopy(teacher_noise_scheduler)     # keeps the 1000-point training grid
        self.disc_update_counter = 0
        self.switch_teacher = False
        if self.discriminator is None:
            logging.warning("No discriminator provided. Adversarial loss will be ignored.")
            self.use_adversarial_loss = False
        else:
            self.use_adversarial_loss = True
        self.disc_backbone = self.teacher_denoiser
        if self.distill_loss_type == "lpips":
            # reference :130-131 `lpips.LPIPS(net="vgg")` — the VGG16 stack on the B200 conv kernels (flash.models.lpips)
            from ..lpips import LPIPS
            if self.vae is None:
                raise ValueError("distill_loss_type='lpips' decodes the latents: a VAE is required (reference :409-410)")
            self.lpips = LPIPS(net="vgg")
        self.K_steps = np.cumsum(self.num_iterations_per_K)
        self.K_prev = self.K[0]
        self.use_cuda_graphs = True
        self.__dict__["_graphed"] = {}
        self.batch_cfg = True

    # ------------------------------------------------------------------ helpers
    def get_sigmas(self, scheduler, timesteps, n_dim=4, dtype=torch.float32, device="cpu"):
        """sigma of each timestep on `scheduler`'s current grid (reference :1043-1060), by exact match."""
        grid = scheduler.timesteps.detach().cpu()
        ts = timesteps.detach().cpu() if torch.is_tensor(timesteps) else torch.as_tensor(timesteps)
        idx = []
        for t in ts.reshape(-1):
            hit = (grid == t).nonzero()
            if hit.numel() != 1:
                raise ValueError(f"timestep {float(t)} matches {hit.numel()} points of the schedule")
            idx.append(int(hit))
        sigma = scheduler.sigmas.detach().cpu()[idx].to(device=device, dtype=dtype).flatten()
        while sigma.dim() < n_dim:
            sigma = sigma.unsqueeze(-1)
        return sigma

    def _prompt_embeddings(self, batch, device):
        if self.pipeline is None:
            missing = [k for k in _EMBED_KEYS if k not in batch]
            if missing:
                raise KeyError(f"no `pipeline` was given, so the batch must carry {missing}")
            return tuple(batch[k].to(device) for k in _EMBED_KEYS)
        self.pipeline.to(device)
        with torch.no_grad():
            out = self.pipeline.encode_prompt(
                prompt=batch["text"], prompt_2=batch["text"], prompt_3=batch["text"],
                negative_prompt=NEGATIVE_PROMPT, negative_prompt_2=NEGATIVE_PROMPT, negative_prompt_3=NEGATIVE_PROMPT,
                do_classifier_free_guidance=True, prompt_embeds=None, negative_prompt_embeds=None,
                pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, clip_skip=False, device=device)
        if self.cpu_offload:
            self.pipeline.to("cpu")
        return out

    def _conditionings(self, batch, device):
        pe, npe, ppe, nppe = self._prompt_embeddings(batch, device)
        return ({"cond": {"vector": ppe, "crossattn": pe}}, {"cond": {"vector": nppe, "crossattn": npe}})

    # ------------------------------------------------------------------ forward (reference :187-371)
    def forward(self, batch: Dict[str, Any], batch_idx=0, step=0, draws: Optional[Dict[str, Any]] = None,
                *args, **kwargs):
        draws = draws or {}
        self.iter_steps += 1
        z = self._encode_inputs(batch) if self.vae is not None else batch[self.input_key]
        conditioning, unconditional_conditioning = self._conditionings(batch, z.device)
        student_conditioning = conditioning

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
        start_idx = int(start_idx)
        sigma0 = float(sched.sigmas[start_idx])                    # host scalar: same value for the whole batch
        if start_idx == 0:
            noisy_sample_init = noise * getattr(sched, "init_noise_sigma", 1.0)
        else:
            noisy_sample_init = sigma0 * noise + (1.0 - sigma0) * z

        if "guidance" in draws:
            guidance_host = float(draws["guidance"])
        else:
            guidance_host = float(torch.rand(1)) * (g_max - g_min) + g_min

        # frozen-teacher rollout enqueued first (graph replays), student forward after it: independent work, same values
        teacher_output = self._teacher_rollout(noisy_sample_init, conditioning, unconditional_conditioning,
                                               start_idx, guidance_host)
        student_v = self.student_denoiser(sample=noisy_sample_init, timestep=start_timestep,
                                          conditioning=student_conditioning)
        student_output = noisy_sample_init - student_v * sigma0

        loss = self._distill_loss(student_output, teacher_output) * self.distill_loss_scale[K_step]
        if self.use_dmd_loss:
            loss = loss + self._dmd_loss(student_output, student_conditioning, conditioning,
                                         unconditional_conditioning, K, K_step, draws) * self.dmd_loss_scale[K_step]
        out = {"teacher_output": teacher_output, "student_output": student_output, "noisy_sample": noisy_sample_init,
               "start_timestep": float(sched.timesteps[start_idx])}
        if self.use_adversarial_loss:
            gan_loss = self._gan_loss(z, batch, student_output, teacher_output, conditioning, step=step, draws=draws)
            loss = loss + self.adversarial_loss_scale[K_step] * gan_loss[0]
            out["loss"] = [loss, gan_loss[1]]
        else:
            out["loss"] = loss.mean()
        return out

    @torch.no_grad()
    def _teacher_rollout(self, noisy_sample_init, conditioning, unconditional_conditioning, start_idx, guidance_scale):
        sched = self.teacher_noise_scheduler
        x = noisy_sample_init.clone().detach()
        B = x.shape[0]
        fused = x.is_cuda and hasattr(sched, "fused_cfg_step")
        w = float(guidance_scale)
        if fused:
            x = x.float().contiguous()
            scratch = torch.zeros_like(x)
        for t in sched.timesteps[start_idx:]:
            timestep = torch.full((B,), float(t), device=x.device)
            v_c, v_u = self._teacher_pair(self.teacher_denoiser, x, timestep, conditioning,
                                          unconditional_conditioning, clone=not fused)
            if fused:
                sched.fused_cfg_step(v_c.contiguous(), v_u.contiguous(), w, t, x, scratch)
            else:
                x = sched.step(w * v_c + (1 - w) * v_u, t, x, return_dict=False)[0]
        return x

    # ------------------------------------------------------------------ losses (reference :372-658)
    def _distill_loss(self, student_output, teacher_output):
        if self.distill_loss_type == "l2":
            return torch.mean(((student_output - teacher_output) ** 2).reshape(student_output.shape[0], -1), 1).mean()
        if self.distill_loss_type == "l1":
            return torch.mean(torch.abs(student_output - teacher_output).reshape(student_output.shape[0], -1), 1).mean()
        if self.distill_loss_type == "lpips":
            # reference :391-411 — centre crop of at most 64x64 latents (clamped to the latent size, unlike the
            # epsilon model's), decode both, clamp to [-1, 1], LPIPS-VGG, mean
            H, W = student_output.shape[2:]
            ch, cw = max((H - 64) // 2, 0), max((W - 64) // 2, 0)
            s_crop = student_output[:, :, ch:min(ch + 64, H), cw:min(cw + 64, W)]
            t_crop = teacher_output[:, :, ch:min(ch + 64, H), cw:min(cw + 64, W)]
            decoded_student = self.vae.decode(s_crop).clamp(-1, 1)
            with torch.no_grad():                  # the teacher output carries no graph in the reference either
                decoded_teacher = self.vae.decode(t_crop).clamp(-1, 1)
            return self.lpips(decoded_student, decoded_teacher).mean()
        raise NotImplementedError(f"Loss type {self.distill_loss_type} not implemented")

    def _dmd_loss(self, student_output, student_conditioning, conditioning, unconditional_conditioning, K, K_step,
                  draws=None):
        draws = draws or {}
        grid = self.teacher_noise_scheduler_copy
        dev = student_output.device
        noise = draws["dmd_noise"] if "dmd_noise" in draws else torch.randn_like(student_output)
        if "dmd_index" in draws:
            index = torch.as_tensor(draws["dmd_index"], device="cpu").long()
        else:
            index = torch.randint(0, self.teacher_noise_scheduler.config.num_train_timesteps,
                                  (student_output.shape[0],), device="cpu")
        timestep = grid.timesteps[index].to(dev)
        sigmas = grid.sigmas[index].to(device=dev, dtype=torch.float32).view(-1, 1, 1, 1)
        noisy_student = sigmas * noise + (1.0 - sigmas) * student_output
        with torch.no_grad():
            real_c, real_u = self._teacher_pair(self.teacher_denoiser, noisy_student, timestep, conditioning,
                                                unconditional_conditioning)
            fake = self.student_denoiser(sample=noisy_student, timestep=timestep, conditioning=student_conditioning)
            if "dmd_guidance" in draws:
                w = float(draws["dmd_guidance"])
            else:
                w = float(torch.rand(1)) * (self.guidance_scale_max[K_step] - self.guidance_scale_min[K_step]) \
                    + self.guidance_scale_min[K_step]
        real = w * real_c + (1 - w) * real_u
        coeff = (-fake) - (-real)                                  # score_fake - score_real, unscaled (:484-488)
        weight = 1.0 / ((student_output - real).abs().mean([1, 2, 3], keepdim=True) + 1e-5).detach()
        return F.mse_loss(student_output, (student_output - weight * coeff).detach(), reduction="mean")

    def _gan_loss(self, z, batch, student_output, teacher_output, conditioning, step=0, draws=None):
        draws = draws or {}
        self.disc_update_counter += 1
        grid = self.teacher_noise_scheduler_copy
        dev = student_output.device
        B = student_output.shape[0]
        noise = draws["gan_noise"] if "gan_noise" in draws else torch.randn_like(student_output)
        real = teacher_output if self.use_teacher_as_real else z
        slots = torch.tensor([-10, -250, -500, -750])              # positions on the training grid (:523-528)
        if "gan_choice" in draws:
            choice = torch.as_tensor(draws["gan_choice"], device="cpu").long()
        else:
            choice = torch.tensor([0.25] * 4).multinomial(B, replacement=True)
        index = slots[choice] % len(grid.timesteps)
        timesteps = grid.timesteps[index].to(dev)
        sigmas = grid.sigmas[index].to(device=dev, dtype=torch.float32).view(-1, 1, 1, 1)
        generator_turn = step % 2 == 0
        fake_in = student_output if generator_turn else student_output.detach()
        noisy_fake = sigmas * noise + (1.0 - sigmas) * fake_in
        noisy_real = sigmas * noise + (1.0 - sigmas) * real
        noisy_sample = torch.cat([noisy_fake, noisy_real], dim=0)
        cond2 = None
        if conditioning is not None:
            cond2 = {"cond": {k: torch.cat([v, v], dim=0) for k, v in conditioning["cond"].items()}}
        # the reference detaches the fake features on the discriminator turn (:597-600): no graph needed there
        with torch.set_grad_enabled(generator_turn and torch.is_grad_enabled()):
            feats = self._call_frozen(self.disc_backbone, noisy_sample, torch.cat([timesteps, timesteps], dim=0),
                                      cond2, return_post_mid_blocks=True)
        f_fake, f_real = feats.chunk(2, dim=0)
        return self._gan_objective(f_fake, f_real, B, generator_turn)

    # ------------------------------------------------------------------ few-step sampler (reference :694-838)
    @torch.no_grad()
    def log_samples(self, batch: Dict[str, Any], input_shape=None, guidance_scale: float = 1.0,
                    teacher_guidance_scale: float = 5.0, max_samples: int = 8, num_steps=20, device="cpu",
                    log_teacher_samples=False, conditioner_inputs: Dict = None, conditioner_uncond_inputs: Dict = None,
                    **sample_kwargs):
        """reference flash_sd3/flash_diffusion_model.py:845-941 (no adapter argument in the SD3 twin)"""
        return self._log_samples(batch, input_shape, guidance_scale, teacher_guidance_scale, max_samples, num_steps,
                                 device, log_teacher_samples, conditioner_inputs, conditioner_uncond_inputs,
                                 **sample_kwargs)

    def sample(self, z, num_steps=20, guidance_scale=1.0, teacher_guidance_scale=5.0, conditioner_inputs=None,
               uncond_conditioner_inputs=None, max_samples=None, verbose=False, log_teacher_samples=False,
               generator=None):
        self.teacher_noise_scheduler.set_timesteps(num_steps)
        self.sampling_noise_scheduler.set_timesteps(num_steps)
        sample = z
        conditioning, unconditional = self._conditionings(conditioner_inputs, z.device)
        if max_samples is not None:
            sample = sample[:max_samples]
            conditioning["cond"] = {k: v[:max_samples] for k, v in conditioning["cond"].items()}
            unconditional["cond"] = {k: v[:max_samples] for k, v in unconditional["cond"].items()}
        sample_init = sample

        def run(denoiser, sched, x, w):
            x = x * getattr(sched, "init_noise_sigma", 1.0)
            for t in sched.timesteps:
                ts = torch.full((x.shape[0],), float(t), device=z.device)
                if w == 1.0:           # guidance 1 multiplies the unconditional branch by exactly 0 (:760-764)
                    v = self._call_frozen(denoiser, x, ts, conditioning)
                else:
                    v_c, v_u = self._teacher_pair(denoiser, x, ts, conditioning, unconditional)
                    v = w * v_c + (1 - w) * v_u
                kw = {"generator": generator} if generator is not None else {}
                x = sched.step(v, t, x, return_dict=False, **kw)[0]
            return x

        out = run(self.student_denoiser, self.sampling_noise_scheduler, sample, float(guidance_scale))
        decoded = self.vae.decode(out) if self.vae is not None else out
        decoded_ref = None
        if log_teacher_samples:
            self.teacher_sampling_noise_scheduler.set_timesteps(num_steps)
            ref = run(self.teacher_denoiser, self.teacher_sampling_noise_scheduler, sample_init,
                      float(teacher_guidance_scale))
            decoded_ref = self.vae.decode(ref) if self.vae is not None else ref
        return decoded, decoded_ref
