"""Builders for the BASELINE configurations on synthetic data (SURVEY.md §8d, Appendix C recipe).

`build_sdxl_distillation(...)` assembles config 2 exactly as examples/train_flash_sdxl.py does — teacher
UNet (:66-118), deep-copied student + LoRA r=64 on to_q/to_k/to_v/to_out.0 (:206-217), discriminator
(:239-267), DPM-Solver++ teacher scheduler with trailing spacing (:221-236), FlashDiffusion (:271-300) and the
two-optimizer TrainingPipeline (:397-412) — with the two things that cannot exist offline replaced:
random-init weights instead of the HF checkpoint, and synthetic text embeddings fed through the reference's
own conditioner API (TorchNNEmbedder(Identity) / TimestepsEmbedder) instead of the CLIP encoders; `vae=None`
so the batch carries latents (flash_diffusion_model.py:182-185) and the distill loss is l2.
"""
import copy
import math

import torch
import torch.nn as nn

from .models.embedders import (ConditionerWrapper, TimestepsEmbedder, TimestepsEmbedderConfig, TorchNNEmbedder,
                               TorchNNEmbedderConfig)
from .models.flash import FlashDiffusion, FlashDiffusionConfig
from .models.lora import LoraConfig
from .models.unets import DiffusersUNet2DCondWrapper
from .schedulers import DPMSolverMultistepScheduler, LCMScheduler
from .trainer import TrainingConfig, TrainingPipeline

SDXL_UNET_KWARGS = dict(
    in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
    mid_block_type="UNetMidBlock2DCrossAttn", up_block_types=["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"],
    only_cross_attention=False, block_out_channels=[320, 640, 1280], layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, dropout=0.0, act_fn="silu", norm_num_groups=32, norm_eps=1e-05,
    cross_attention_dim=2048, transformer_layers_per_block=[1, 2, 10], attention_head_dim=[5, 10, 20],
    use_linear_projection=True, class_embed_type="projection", projection_class_embeddings_input_dim=2816)

SD15_UNET_KWARGS = dict(          # examples/train_flash_sd.py:56-114
    in_channels=4, out_channels=4,
    down_block_types=["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    up_block_types=["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"],
    block_out_channels=[320, 640, 1280, 1280], layers_per_block=2, cross_attention_dim=768,
    transformer_layers_per_block=1, attention_head_dim=8, use_linear_projection=True, class_embed_type=None)

TINY_UNET_KWARGS = dict(
    in_channels=4, out_channels=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"],
    up_block_types=["CrossAttnUpBlock2D", "UpBlock2D"], block_out_channels=[64, 128], layers_per_block=1,
    cross_attention_dim=96, transformer_layers_per_block=[1, 2], attention_head_dim=[1, 2],
    use_linear_projection=True, class_embed_type="projection", projection_class_embeddings_input_dim=48 + 3 * 16)


@torch.no_grad()
def init_random_(module: nn.Module, seed: int):
    """Deterministic random init directly on the module's device (fan-in scaled normal weights, zero biases,
    unit norms).  Same seed -> same weights on every rank."""
    g = torch.Generator(device=next(module.parameters()).device).manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() >= 2:
            fan_in = p[0].numel()
            p.normal_(0.0, 1.0 / math.sqrt(fan_in), generator=g)
        elif name.endswith("bias"):
            p.zero_()
        else:
            p.fill_(1.0)           # GroupNorm / LayerNorm weights
    return module


def sdxl_discriminator(color_dim=1280, feature_dim=256):
    """examples/train_flash_sdxl.py:239-267"""
    d = feature_dim
    return nn.Sequential(
        nn.Conv2d(color_dim, d, 4, 2, 1, bias=False), nn.SiLU(True),
        nn.Conv2d(d, d * 2, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 2), nn.SiLU(True),
        nn.Conv2d(d * 2, d * 4, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 4), nn.SiLU(True),
        nn.Conv2d(d * 4, 1, 4, 1, 0, bias=False), nn.Flatten())


def tiny_discriminator(color_dim=128):
    return nn.Sequential(nn.Conv2d(color_dim, 32, 4, 2, 1, bias=False), nn.SiLU(True),
                         nn.Conv2d(32, 32, 4, 2, 1, bias=False), nn.GroupNorm(4, 32), nn.SiLU(True),
                         nn.Conv2d(32, 1, 4, 1, 0, bias=False), nn.Flatten())


def text_only_conditioner():
    """SD1.5: cross-attention text embedding only (examples/train_flash_sd.py: one CLIP embedder, no vector)."""
    return ConditionerWrapper([TorchNNEmbedder(TorchNNEmbedderConfig(
        input_key="text_emb", nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}]))])


def synthetic_conditioner(fourier_channels=256):
    """text / pooled embeddings through Identity embedders, size/crop ids through
    TimestepsEmbedder — the same slots the CLIP embedders fill in examples/train_flash_sdxl.py:137-195."""
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}])
    return ConditionerWrapper([
        TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
        TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="original_size_as_tuple", num_channels=fourier_channels)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="crop_coords_top_left", num_channels=fourier_channels)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="target_size_as_tuple", num_channels=fourier_channels)),
    ])


def synthetic_batch(B, latent_hw, ctx_tokens, ctx_dim, pooled_dim, seed, device="cpu", pin=False, image_px=1024.0):
    g = torch.Generator().manual_seed(seed)
    batch = {
        "image": torch.randn(B, 4, latent_hw, latent_hw, generator=g),
        "text_emb": torch.randn(B, ctx_tokens, ctx_dim, generator=g),
        "pooled_emb": torch.randn(B, pooled_dim, generator=g),
        "original_size_as_tuple": torch.tensor([[image_px, image_px]] * B),
        "crop_coords_top_left": torch.zeros(B, 2),
        "target_size_as_tuple": torch.tensor([[image_px, image_px]] * B),
    }
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def build_distillation(unet_kwargs, discriminator, device, *, lora_rank=64, K=32, seed=1234, lora_b_std=0.01,
                       fourier_channels=256, stage=2, lr=1e-5, conditioner=None, ucg_keys=("text_emb", "pooled_emb")):
    """teacher / student(LoRA) / discriminator / FlashDiffusion / TrainingPipeline on `device`."""
    with torch.device("meta"):
        teacher = DiffusersUNet2DCondWrapper(**unet_kwargs)
    teacher = teacher.to_empty(device=device)
    init_random_(teacher, seed)
    student = copy.deepcopy(teacher)
    torch.manual_seed(seed + 3)          # LoRA A draws from the global generator: pin it to the recipe's seed
    student.add_adapter(LoraConfig(r=lora_rank, lora_alpha=lora_rank, init_lora_weights="gaussian",
                                   target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    if lora_b_std > 0:
        g = torch.Generator(device=device).manual_seed(seed + 1)
        with torch.no_grad():
            for n, p in student.named_parameters():
                if "lora_B" in n:
                    p.normal_(0.0, lora_b_std, generator=g)
    teacher.freeze()
    torch.manual_seed(seed + 2)
    discriminator = discriminator.to(device)
    # flash_sdxl.yaml:11-32 — K=32 every stage, mixture over 4 modes; stage index selects the yaml's per-stage
    # loss scales and mode probabilities (stage 2 = uniform modes, E[n] = 20)
    adv = [0.0, 0.1, 0.2, 0.3][stage]
    dmd = [0.0, 0.3, 0.5, 0.7][stage]
    probs = [[0.0, 0.0, 0.5, 0.5], [0.1, 0.3, 0.3, 0.3], [0.25, 0.25, 0.25, 0.25], [0.4, 0.2, 0.2, 0.2]][stage]
    cfg = FlashDiffusionConfig(
        K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=3.0, guidance_scale_max=13.0,
        distill_loss_type="l2", ucg_keys=list(ucg_keys), timestep_distribution="mixture",
        mixture_num_components=4, mixture_var=0.5, use_dmd_loss=True, dmd_loss_scale=dmd, distill_loss_scale=1.0,
        adversarial_loss_scale=adv, gan_loss_type="lsgan", mode_probs=[probs], use_teacher_as_real=False,
        use_empty_prompt=False, input_key="image")
    sched = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0",
                                                        subfolder="scheduler", timestep_spacing="trailing")
    lcm = LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                       timestep_spacing="trailing")
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                           sampling_noise_scheduler=lcm, vae=None,
                           conditioner=conditioner if conditioner is not None else synthetic_conditioner(fourier_channels),
                           discriminator=discriminator).to(device)
    pipe = TrainingPipeline(model, TrainingConfig(
        optimizers_name=["AdamW", "AdamW"], learning_rates=[lr, lr],
        trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    return model, pipe


def build_sdxl_distillation(device, **kw):
    return build_distillation(SDXL_UNET_KWARGS, sdxl_discriminator(), device, **kw)


def sd15_discriminator(color_dim=1280, d=64):
    """examples/train_flash_sd.py:221-240 (mid-block features 1280 x 8 x 8 -> 1 logit)."""
    return nn.Sequential(nn.Conv2d(color_dim, d, 3, 1, 1), nn.SiLU(True),
                         nn.Conv2d(d, d * 2, 4, 2, 1, bias=False), nn.SiLU(True), nn.GroupNorm(4, d * 2),
                         nn.Conv2d(d * 2, 1, 4, 1, 0, bias=False), nn.Flatten())


def build_sd15_distillation(device, **kw):
    """Config 1 objects (examples/train_flash_sd.py + configs/flash_sd.yaml: SD1.5 UNet, LoRA r=128 on q/k/v/out,
    K=32).  Head dims 40 / 80 / 160 run through zero-padded packs and the generic attention kernels."""
    kw.setdefault("lora_rank", 128)
    return build_distillation(SD15_UNET_KWARGS, sd15_discriminator(), device, conditioner=text_only_conditioner(),
                              ucg_keys=("text_emb",), **kw)


PIXART_KWARGS = dict(   # examples/train_flash_pixart.py:65-86
    sample_size=128, num_layers=28, attention_head_dim=72, in_channels=4, out_channels=8, patch_size=2,
    attention_bias=True, num_attention_heads=16, cross_attention_dim=1152, activation_fn="gelu-approximate",
    num_embeds_ada_norm=1000, norm_type="ada_norm_single", norm_elementwise_affine=False, norm_eps=1e-6,
    caption_channels=4096, projection_class_embeddings_input_dim=256, time_embed_dim=1152,
    timesteps_embedding_num_channels=256, use_concat_vector_conditioning=True, num_vector_conditionings=3)


# examples/train_flash_pixart.py:239-252 and examples/train_flash_sd3.py:104-117 (the same list)
DIT_LORA_TARGETS = ["to_q", "to_k", "to_v", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2", "proj",
                    "linear", "linear_1", "linear_2"]


def pixart_discriminator(color_dim=4, d=64):
    """examples/train_flash_pixart.py:277-325 (on the 4-channel backbone output: the DiT wrapper has no mid-block
    features to return)."""
    return nn.Sequential(nn.Conv2d(color_dim, d, 4, 2, 1, bias=False), nn.SiLU(True),
                         nn.Conv2d(d, d * 2, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 2), nn.SiLU(True),
                         nn.Conv2d(d * 2, d * 4, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 4), nn.SiLU(True),
                         nn.Conv2d(d * 4, d * 8, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 8), nn.SiLU(True),
                         nn.Conv2d(d * 8, d * 16, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 16), nn.SiLU(True),
                         nn.Conv2d(d * 16, 1, 4, 1, 0, bias=False), nn.Flatten())


def build_pixart_distillation(device, lora_rank=64, seed=1234, K=16, lr=1e-5, kwargs=None, discriminator=None,
                              lora_b_std=0.0):
    """Config 3 objects (examples/train_flash_pixart.py + configs/flash_pixart.yaml): PixArt-alpha XL/2 teacher, LoRA
    r=64 student (peft default lora_alpha 8) on the script's Linear targets, DPM-Solver++ trailing K=16 teacher, LCM
    sampler, DMD + lsgan through the frozen backbone.  Returns (model, pipe)."""
    from .models.embedders import PrecomputedTextEmbedder, PrecomputedTextEmbedderConfig
    from .models.transformers import DiffusersTransformer2DWrapper
    from .models.transformers.transformers import sincos_2d
    kwargs = kwargs or PIXART_KWARGS
    with torch.device("meta"):
        teacher = DiffusersTransformer2DWrapper(**kwargs)
    teacher = teacher.to_empty(device=device)
    init_random_(teacher, seed)
    grid = kwargs["sample_size"] // kwargs["patch_size"]
    D = kwargs["num_attention_heads"] * kwargs["attention_head_dim"]
    teacher.pos_embed.pos_embed = torch.from_numpy(
        sincos_2d(D, grid, grid, max(kwargs["sample_size"] // 64, 1))).float()[None].to(device)
    student = copy.deepcopy(teacher)
    student.add_adapter(LoraConfig(r=lora_rank, lora_alpha=8, target_modules=DIT_LORA_TARGETS))
    _perturb_lora_b(student, lora_b_std, seed + 1)
    teacher.freeze()
    fc = kwargs["projection_class_embeddings_input_dim"]
    conditioner = ConditionerWrapper([
        PrecomputedTextEmbedder(PrecomputedTextEmbedderConfig(input_key="text_emb", mask_key="text_mask")),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="resolution", num_channels=fc)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="aspect_ratio", num_channels=fc))])
    cfg = FlashDiffusionConfig(K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=2.0, guidance_scale_max=9.0,
                               distill_loss_type="l2", ucg_keys=["text_emb"], use_dmd_loss=True, gan_loss_type="lsgan",
                               timestep_distribution="mixture", mixture_num_components=4, mixture_var=0.5,
                               input_key="image")
    name = "PixArt-alpha/PixArt-XL-2-1024-MS"
    sched = DPMSolverMultistepScheduler.from_pretrained(name, subfolder="scheduler", timestep_spacing="trailing")
    lcm = LCMScheduler.from_pretrained(name, subfolder="scheduler", timestep_spacing="trailing")
    disc = discriminator if discriminator is not None else pixart_discriminator(kwargs["in_channels"])
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                           sampling_noise_scheduler=lcm, vae=None, conditioner=conditioner, discriminator=disc).to(device)
    pipe = TrainingPipeline(model, TrainingConfig(
        optimizers_name=["AdamW", "AdamW"], learning_rates=[lr, lr],
        trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    return model, pipe


def build_pixart_sampler(device, lora_rank=64, seed=1234):
    """PixArt-alpha LoRA student inside a FlashDiffusion for the few-step sampler (config 5)."""
    return build_pixart_distillation(device, lora_rank=lora_rank, seed=seed)[0]


def _perturb_lora_b(module, std, seed):
    """peft initialises lora_B to zero (student == teacher); a small non-zero B makes parity / timing runs exercise the
    adapter products."""
    if std > 0:
        g = torch.Generator(device=next(module.parameters()).device).manual_seed(seed)
        with torch.no_grad():
            for n, p in module.named_parameters():
                if "lora_B" in n:
                    p.normal_(0.0, std, generator=g)


def pixart_batch(B, seed, device, tokens=120, valid=77, hw=128, ctx_dim=4096):
    g = torch.Generator().manual_seed(seed)
    batch = {"image": torch.randn(B, 4, hw, hw, generator=g), "text_emb": torch.randn(B, tokens, ctx_dim, generator=g),
             "text_mask": (torch.arange(tokens)[None] < valid).long().repeat(B, 1),
             "resolution": torch.tensor([[1024., 1024.]] * B), "aspect_ratio": torch.tensor([[1.0]] * B)}
    return {k: v.to(device) for k, v in batch.items()}


def build_tiny_distillation(device, **kw):
    kw.setdefault("lora_rank", 64)
    kw.setdefault("K", 4)
    kw.setdefault("fourier_channels", 8)
    return build_distillation(TINY_UNET_KWARGS, tiny_discriminator(), device, **kw)


SD3_KWARGS = dict(   # examples/train_flash_sd3.py:65-77
    sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64, num_attention_heads=24,
    joint_attention_dim=4096, caption_projection_dim=1536, pooled_projection_dim=2048, out_channels=16,
    pos_embed_max_size=192)

SD3_LORA_TARGETS = DIT_LORA_TARGETS          # examples/train_flash_sd3.py:104-117


def sd3_discriminator(color_dim=16, d=64):
    """examples/train_flash_sd3.py:143-181 (on the 16-channel backbone output)."""
    return nn.Sequential(nn.Conv2d(color_dim, d, 4, 2, 1, bias=False), nn.SiLU(True),
                         nn.Conv2d(d, d * 2, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 2), nn.SiLU(True),
                         nn.Conv2d(d * 2, d * 4, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 4), nn.SiLU(True),
                         nn.Conv2d(d * 4, d * 8, 4, 2, 1, bias=False), nn.GroupNorm(4, d * 8), nn.SiLU(True),
                         nn.Conv2d(d * 8, 1, 4, 1, 0, bias=False), nn.Flatten())


def build_sd3(device, kwargs=None, lora_rank=64, seed=1234, K=32, discriminator=None, lora_b_std=0.0):
    """Config 4 objects (examples/train_flash_sd3.py + configs/flash_sd3.yaml): SD3-medium MMDiT teacher, LoRA student
    (lora_alpha = peft default 8) on the script's Linear targets, flow-matching schedulers, DMD + lsgan, inside a
    `FlashDiffusionSD3`."""
    from .models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    from .models.transformers import DiffusersSD3Transformer2DWrapper
    from .schedulers import FlashFlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteScheduler
    kwargs = kwargs or SD3_KWARGS
    with torch.device("meta"):
        teacher = DiffusersSD3Transformer2DWrapper(**kwargs)
    teacher = teacher.to_empty(device=device)
    init_random_(teacher, seed)
    pe = teacher.pos_embed
    fresh = type(pe)(kwargs["sample_size"], kwargs["patch_size"], kwargs["in_channels"], teacher.inner_dim,
                     kwargs["pos_embed_max_size"])
    pe.pos_embed = fresh.pos_embed.to(device)
    student = copy.deepcopy(teacher)
    if lora_rank:
        student.add_adapter(LoraConfig(r=lora_rank, lora_alpha=8, target_modules=SD3_LORA_TARGETS))
        _perturb_lora_b(student, lora_b_std, seed + 1)
    teacher.freeze()
    cfg = FlashDiffusionSD3Config(K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=7.0,
                                  guidance_scale_max=13.0, distill_loss_type="l2", use_dmd_loss=True,
                                  gan_loss_type="lsgan", input_key="image")
    mk = lambda cls, **kw: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", **kw)
    return FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                             teacher_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler, timestep_spacing="trailing"),
                             sampling_noise_scheduler=mk(FlashFlowMatchEulerDiscreteScheduler, timestep_spacing="trailing"),
                             teacher_sampling_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler),
                             discriminator=discriminator if discriminator is not None
                             else sd3_discriminator(kwargs["out_channels"])).to(device)


def build_sd3_distillation(device, lr=1e-5, **kw):
    """(model, pipe) for the SD3 distillation step; two AdamW optimizers as in examples/train_flash_sd3.py:300-330."""
    model = build_sd3(device, **kw)
    pipe = TrainingPipeline(model, TrainingConfig(
        optimizers_name=["AdamW", "AdamW"], learning_rates=[lr, lr],
        trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    return model, pipe


def sd3_batch(B, seed, device, kwargs=None, tokens=154, hw=None):
    kwargs = kwargs or SD3_KWARGS
    hw = hw or kwargs["sample_size"]
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    j, p = kwargs["joint_attention_dim"], kwargs["pooled_projection_dim"]
    batch = {"image": r(B, kwargs["in_channels"], hw, hw), "prompt_embeds": r(B, tokens, j),
             "negative_prompt_embeds": r(B, tokens, j), "pooled_prompt_embeds": r(B, p),
             "negative_pooled_prompt_embeds": r(B, p)}
    return {k: v.to(device) for k, v in batch.items()}
