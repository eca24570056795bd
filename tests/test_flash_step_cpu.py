"""CPU tests of the host logic: FlashDiffusion / TrainingPipeline / schedulers / conditioners.

The denoisers plugged in here are the fp32 ORACLE UNets (tests may use oracle/): the product denoiser is
CUDA-only by design.  Properties reproduced from the reference's own tests
(tests/test_flash/test_flash_diffusion.py:146-222): loss sign pattern per `step`, who-gets-updated invariants.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from flash.models.embedders import (ConditionerWrapper, TimestepsEmbedder, TimestepsEmbedderConfig,
                                    TorchNNEmbedder, TorchNNEmbedderConfig)
from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
from flash.schedulers import DDPMScheduler, DPMSolverMultistepScheduler, LCMScheduler
from flash.trainer import TrainingConfig, TrainingPipeline
from oracle import flash_step as OF
from oracle import schedulers as OS
from oracle.unet import LoraConfig, UNet2DConditionOracle

TINY = dict(in_channels=4, out_channels=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"],
            up_block_types=["CrossAttnUpBlock2D", "UpBlock2D"], block_out_channels=[32, 64], layers_per_block=1,
            cross_attention_dim=32, transformer_layers_per_block=1, attention_head_dim=[1, 2], norm_num_groups=8,
            use_linear_projection=True, class_embed_type="projection", projection_class_embeddings_input_dim=16 + 3 * 8)


def _conditioner():
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}])
    return ConditionerWrapper([
        TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
        TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="original_size_as_tuple", num_channels=4)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="crop_coords_top_left", num_channels=4)),
        TimestepsEmbedder(TimestepsEmbedderConfig(input_key="target_size_as_tuple", num_channels=4, input_dim=7)),
    ])


def _batch(B=2, hw=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"image": torch.randn(B, 4, hw, hw, generator=g), "text_emb": torch.randn(B, 5, 32, generator=g),
            "pooled_emb": torch.randn(B, 16, generator=g),
            "original_size_as_tuple": torch.tensor([[1024., 1024.]] * B), "crop_coords_top_left": torch.zeros(B, 2),
            "target_size_as_tuple": torch.tensor([[1024., 1024.]] * B)}


def _model(gan="lsgan", K=4, dmd=True, seed=0):
    torch.manual_seed(seed)
    teacher = UNet2DConditionOracle(**TINY)
    student = copy.deepcopy(teacher)
    student.add_adapter(LoraConfig(r=8, lora_alpha=8, init_lora_weights="gaussian",
                                   target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    for n, p in student.named_parameters():
        if "lora_B" in n:
            nn.init.normal_(p, std=0.05)
    teacher.freeze()
    disc = nn.Sequential(nn.Conv2d(64, 16, 4, 2, 1, bias=False), nn.SiLU(True), nn.Conv2d(16, 1, 4, 1, 0, bias=False),
                         nn.Flatten())
    cfg = FlashDiffusionConfig(K=[K], num_iterations_per_K=[100], guidance_scale_min=3.0, guidance_scale_max=7.0,
                               distill_loss_type="l2", ucg_keys=["text_emb", "pooled_emb"], use_dmd_loss=dmd,
                               gan_loss_type=gan, timestep_distribution="mixture", mixture_num_components=2,
                               mixture_var=0.5, switch_teacher=False, allow_full_noise=False)
    sched = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler", timestep_spacing="trailing")
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                           sampling_noise_scheduler=LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing"),
                           vae=None, conditioner=_conditioner(), discriminator=disc)
    return model


def test_config_ignores_unknown_kwargs_and_broadcasts():
    c = FlashDiffusionConfig(K=[8, 8], num_iterations_per_K=[1, 2], removed_field=3)
    assert c.guidance_scale_min == [3.0, 3.0] and c.mode_probs == [[0.25] * 4] * 2
    with pytest.raises(Exception, match="Number of timesteps must match"):
        FlashDiffusionConfig(K=[8, 8], num_iterations_per_K=[1])


def test_conditioner_wrapper_ucg_semantics():
    cw, b = _conditioner(), _batch()
    c = cw(b)["cond"]
    assert c["crossattn"].shape == (2, 5, 32) and c["vector"].shape == (2, 16 + 24)
    u = cw(b, ucg_keys=["text_emb", "pooled_emb"])["cond"]
    assert (u["crossattn"] == 0).all() and (u["vector"][:, :16] == 0).all()
    assert torch.equal(u["vector"][:, 16:], c["vector"][:, 16:])       # size/crop features kept


def test_dpm_scheduler_matches_oracle_rollout():
    sched = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    sched.set_timesteps(32)
    assert sched.timesteps.tolist() == OS.trailing_timesteps(32).tolist()
    assert sched.timesteps[0] == 999 and sched.timesteps[-1] == 30
    ac = OS.alphas_cumprod()
    assert np.allclose(sched.alphas_cumprod.numpy(), ac, rtol=1e-5)
    torch.manual_seed(0)
    x0 = torch.randn(2, 4, 8, 8, dtype=torch.float64)
    W = torch.randn(4, 4, dtype=torch.float64) * 0.3

    def eps_fn(x, t):
        return torch.tanh(torch.einsum("ij,bjhw->bihw", W, x)) * (1 + t / 1000.0)

    for K, start in [(32, 0), (32, 8), (32, 24), (8, 3), (4, 0)]:
        ref = OS.dpm_rollout(eps_fn, x0.clone(), ac, K, start)
        sched.set_timesteps(K)
        x = x0.clone()
        for t in sched.timesteps[start:]:
            x = sched.step(eps_fn(x, int(t)), t, x)[0]
        assert torch.allclose(x, ref, rtol=1e-4, atol=1e-4), (K, start)


def test_lcm_scheduler_timesteps_and_scalings():
    lcm = LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0")
    lcm.set_timesteps(4)
    assert lcm.timesteps.tolist() == [999, 759, 499, 259] == OS.lcm_timesteps(4).tolist()
    cs, co = lcm.scalings(500)
    rcs, rco = OS.lcm_scalings(500.0)
    assert abs(cs - rcs) < 1e-12 and abs(co - rco) < 1e-12


def _draws(B=2, hw=16, seed=3, start_idx=1):
    g = torch.Generator().manual_seed(seed)
    return dict(noise=torch.randn(B, 4, hw, hw, generator=g), start_idx=start_idx, guidance=5.5,
                dmd_noise=torch.randn(B, 4, hw, hw, generator=g), dmd_timestep=torch.tensor([700, 120, 333, 901][:B]),
                dmd_guidance=4.25, gan_noise=torch.randn(B, 4, hw, hw, generator=g),
                gan_timesteps=torch.tensor([250, 750, 10, 500][:B]))


@pytest.mark.parametrize("gan", ["lsgan", "hinge", "vanilla"])
@pytest.mark.parametrize("step", [0, 1])
def test_forward_matches_oracle_step(gan, step):
    """Product host logic (batched 2B CFG, scheduler classes, conditioner dedupe) == line-by-line oracle."""
    model = _model(gan=gan)
    b, d = _batch(), _draws()
    out = model(b, step=step, draws=d)
    cw = model.conditioner
    cond, uncond = cw(b, set_ucg_rate_zero=True), cw(b, ucg_keys=["text_emb", "pooled_emb"])
    ref = OF.flash_forward(model.student_denoiser, model.teacher_denoiser, model.discriminator, b["image"], cond,
                           uncond, d, K=4, step=step, gan_loss_type=gan)
    assert torch.allclose(out["student_output"], ref["student_output"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["teacher_output"], ref["teacher_output"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["loss"][0], ref["loss_G"], rtol=1e-4, atol=1e-6)
    if step == 0:
        assert out["loss"][1] == 0 and out["loss"][0] > 0
    else:
        assert torch.allclose(out["loss"][1], ref["loss_D"], rtol=1e-4, atol=1e-6) and out["loss"][1] > 0


def test_start_idx_zero_uses_pure_noise():
    model = _model()
    b, d = _batch(), _draws(start_idx=0)
    out = model(b, step=0, draws=d)
    assert torch.equal(out["noisy_sample"], d["noise"]) and out["start_timestep"] == 999


def test_start_index_pmf_is_mixture():
    model = _model(K=32)
    model.mixture_num_components, model.mode_probs, model.mixture_var = [4], [[0.25] * 4], [0.5]
    pmf = model._start_index_pmf(32, 0)
    assert abs(float(pmf.sum()) - 1) < 1e-6
    assert sorted(torch.topk(pmf, 4).indices.tolist()) == [0, 8, 16, 24]
    assert float(pmf[[0, 8, 16, 24]].sum()) > 0.75


def test_training_step_update_invariants():
    """reference tests/test_flash/test_flash_diffusion.py:155-187"""
    model = _model()
    pipe = TrainingPipeline(model, TrainingConfig(optimizers_name=["AdamW", "AdamW"], learning_rates=[1e-3, 1e-3],
                                                  trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    assert not pipe.automatic_optimization
    assert all("lora_" in n for n, p in model.student_denoiser.named_parameters() if p.requires_grad)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    out = pipe.training_step(_batch(), 0)
    assert out["loss_optimizer_0"] > 0 and out["loss_optimizer_1"] > 0
    changed = {n for n, p in model.named_parameters() if not torch.equal(p, before[n])}
    assert all(("lora_" in n and n.startswith("student_denoiser")) or n.startswith("discriminator") for n in changed)
    assert any(n.startswith("student_denoiser") for n in changed) and any(n.startswith("discriminator") for n in changed)
    assert not any(n.startswith("teacher_denoiser") for n in changed)


def test_sample_runs_four_steps():
    model = _model()
    b = _batch()
    out, ref = model.sample(torch.randn(2, 4, 16, 16), num_steps=4, guidance_scale=1.0, conditioner_inputs=b)
    assert out.shape == (2, 4, 16, 16) and ref is None and torch.isfinite(out).all()


def test_ddpm_scheduler_protocol():
    s = DDPMScheduler()
    s.set_timesteps(10)
    assert len(s.timesteps) == 10 and s.init_noise_sigma == 1.0
    x = torch.randn(1, 4, 8, 8)
    assert s.step(torch.randn_like(x), s.timesteps[0], x)[0].shape == x.shape
    assert s.add_noise(x, torch.randn_like(x), torch.tensor([5])).shape == x.shape


def test_lean_discriminator_turn_is_output_preserving():
    """Eliding the generator objective on the discriminator turn leaves loss[1] and the discriminator
    update unchanged."""
    b, d = _batch(), _draws()
    strict = _model()
    lean = _model()
    lean.elide_unused_generator_pass = True
    o_s = strict(b, step=1, draws=d)
    o_l = lean(b, step=1, draws=d)
    assert o_l["loss"][0] is None and o_l["teacher_output"] is None
    assert torch.allclose(o_l["loss"][1], o_s["loss"][1], rtol=1e-6, atol=1e-8)
    assert torch.allclose(o_l["student_output"], o_s["student_output"], rtol=1e-6, atol=1e-7)
    o_s["loss"][1].backward()
    o_l["loss"][1].backward()
    for (n1, p1), (n2, p2) in zip(strict.discriminator.named_parameters(), lean.discriminator.named_parameters()):
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-5, atol=1e-8), n1
    assert all(p.grad is None for p in lean.student_denoiser.parameters())
    # generator turn untouched
    assert lean(b, step=0, draws=d)["loss"][0] is not None


def test_log_samples_keys_and_sample_cap():
    """reference flash_diffusion_model.py:917-1019: one entry per num_steps (plus the teacher's when asked), N capped by
    max_samples and by the shortest conditioning entry; input_shape is mandatory without a VAE."""
    model = _model()
    model.teacher_sampling_noise_scheduler = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    b = _batch(B=2)
    logs = model.log_samples(dict(b), input_shape=(4, 16, 16), num_steps=[1, 2], max_samples=8, guidance_scale=1.0,
                             teacher_guidance_scale=3.0, log_teacher_samples=True)
    assert set(logs) == {"samples_1_steps/LCMScheduler_1.0_cfg/student", "samples_2_steps/LCMScheduler_1.0_cfg/student",
                         "samples_1_steps/DPMSolverMultistepScheduler_3.0_cfg/teacher",
                         "samples_2_steps/DPMSolverMultistepScheduler_3.0_cfg/teacher"}
    assert all(v.shape == (2, 4, 16, 16) and torch.isfinite(v).all() for v in logs.values())
    one = model.log_samples(dict(b), input_shape=(4, 16, 16), num_steps=1, max_samples=1)
    assert list(one.values())[0].shape[0] == 1
    with pytest.raises(ValueError, match="input_shape"):
        model.log_samples(dict(b), num_steps=1)


# ------------------------------------------------------------------------------------------ round-2 additions
def test_scheduler_configs_are_keyed_by_repo():
    """`from_pretrained` must not hand the SD config to every repo."""
    sd = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    px = DPMSolverMultistepScheduler.from_pretrained("PixArt-alpha/PixArt-XL-2-1024-MS", timestep_spacing="trailing")
    assert sd.config.beta_schedule == "scaled_linear" and abs(float(sd.betas[0]) - 0.00085) < 1e-7
    assert px.config.beta_schedule == "linear" and abs(float(px.betas[0]) - 1e-4) < 1e-8 and abs(float(px.betas[-1]) - 0.02) < 1e-7
    assert px.config.steps_offset == 0 and sd.config.steps_offset == 1
    assert not torch.allclose(sd.alphas_cumprod, px.alphas_cumprod)
    with pytest.raises(ValueError, match="unknown scheduler repo"):
        DPMSolverMultistepScheduler.from_pretrained("somebody/some-model")


def test_dpm_second_order_at_penultimate_step_for_short_schedules():
    """with solver_order 2 upstream takes the 2M step at index n-2 also when n < 15 (`lower_order_second`
    only demotes a third-order solver)."""
    s = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing")
    s.set_timesteps(4)
    coefs = [s.step_coefficients(t) for t in s.timesteps]
    assert coefs[0][4] == 0.0            # first step: no history
    assert coefs[1][4] != 0.0
    assert coefs[2][4] != 0.0            # index n-2 with n = 4 < 15: still second order
    assert coefs[3][4] == 0.0            # final step: first order (final_sigmas_type="zero")


@pytest.mark.parametrize("mode", ["closed_form", "schedule_index"])
def test_add_noise_modes_match_oracle(mode):
    """SURVEY §8c decision (1), both readings: off-schedule DMD/GAN timesteps either use the closed form or (diffusers
    >= 0.27) the sigma at the LAST position of the current K-step table."""
    K = 32
    s = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing",
                                                    add_noise_mode=mode)
    s.set_timesteps(K)
    ac = OS.alphas_cumprod()
    g = torch.Generator().manual_seed(1)
    x, e = torch.randn(5, 4, 8, 8, generator=g), torch.randn(5, 4, 8, 8, generator=g)
    t = torch.tensor([999, 30, 500, 10, 968])         # 999 / 30 / 968 are on the trailing table, 500 / 10 are not
    got = s.add_noise(x, e, t)
    ref = OS.add_noise(ac, x, e, t) if mode == "closed_form" else OS.add_noise_schedule_index(ac, x, e, t, K)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    closed = OS.add_noise(ac, x, e, t)
    on_table = [0, 1, 4]
    assert torch.allclose(got[on_table], closed[on_table], rtol=1e-5, atol=1e-6)     # identical on the table
    if mode == "schedule_index":
        last = OS.add_noise(ac, x, e, torch.full((5,), 30))
        assert torch.allclose(got[[2, 3]], last[[2, 3]], rtol=1e-5, atol=1e-6)      # misses -> last index (t = 30)
        assert not torch.allclose(got[[2, 3]], closed[[2, 3]], atol=1e-3)


@pytest.mark.parametrize("step", [0, 1])
def test_forward_matches_oracle_step_schedule_index_mode(step):
    model = _model()
    model.teacher_noise_scheduler = DPMSolverMultistepScheduler.from_pretrained(
        "stabilityai/stable-diffusion-xl-base-1.0", timestep_spacing="trailing", add_noise_mode="schedule_index")
    b, d = _batch(), _draws()
    out = model(b, step=step, draws=d)
    cw = model.conditioner
    cond, uncond = cw(b, set_ucg_rate_zero=True), cw(b, ucg_keys=["text_emb", "pooled_emb"])
    ref = OF.flash_forward(model.student_denoiser, model.teacher_denoiser, model.discriminator, b["image"], cond,
                           uncond, d, K=4, step=step, add_noise_mode="schedule_index")
    ref_closed = OF.flash_forward(model.student_denoiser, model.teacher_denoiser, model.discriminator, b["image"], cond,
                                  uncond, d, K=4, step=step)
    assert torch.allclose(out["loss"][0], ref["loss_G"], rtol=1e-4, atol=1e-6)
    assert not torch.allclose(ref["loss_G"], ref_closed["loss_G"], rtol=1e-4)       # the two readings do differ
    if step == 1:
        assert torch.allclose(out["loss"][1], ref["loss_D"], rtol=1e-4, atol=1e-6)


def test_pixart_position_table_is_recomputed_for_other_grids():
    """diffusers' PatchEmbed recomputes the 2-D sincos table for (h, w) != configured grid."""
    from flash.models.transformers.transformers import PatchEmbed
    from oracle.dit import PatchEmbed as OraclePatchEmbed
    pe = PatchEmbed(sample_size=32, patch_size=2, in_channels=4, embed_dim=16)
    ope = OraclePatchEmbed(sample_size=32, patch_size=2, in_channels=4, embed_dim=16)
    assert torch.equal(pe.table(16, 16), pe.pos_embed[0])
    t = pe.table(8, 12)
    assert t.shape == (96, 16) and not torch.allclose(t, pe.pos_embed[0, :96])
    with torch.no_grad():
        for p in ope.parameters():
            p.zero_()
        got = ope(torch.zeros(1, 4, 16, 24))[0]        # zero conv -> the table itself
    assert torch.allclose(got, t, atol=1e-6)
