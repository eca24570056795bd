"""CPU tests of the rectified-flow (SD3) objective host logic: FlashDiffusionSD3 + flow-matching schedulers.

The denoisers plugged in are the fp32 ORACLE MMDiT (tests may use oracle/): the product MMDiT wrapper is CUDA-only and
forward-only.  The product class restructures the reference step (one 2B teacher call, host sigma look-ups, fused
update on CUDA); the oracle restates it line by line — both must give the same numbers on the same draws.
"""
import copy

import pytest
import torch
import torch.nn as nn

from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
from flash.models.lora import LoraConfig, inject_lora
from flash.schedulers import FlashFlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteScheduler
from flash.trainer import TrainingConfig, TrainingPipeline
from oracle import flash_step_sd3 as O3
from oracle.sd3 import SD3TransformerOracle

TINY = dict(sample_size=8, patch_size=2, in_channels=4, num_layers=2, attention_head_dim=8, num_attention_heads=2,
            joint_attention_dim=12, caption_projection_dim=16, pooled_projection_dim=10, out_channels=4,
            pos_embed_max_size=8)


def _denoisers(seed=0):
    torch.manual_seed(seed)
    teacher = SD3TransformerOracle(**TINY)
    with torch.no_grad():
        for p in teacher.parameters():
            if p.dim() >= 2:
                p.normal_(0, 1.0 / p[0].numel() ** 0.5)
            else:
                p.normal_(0, 0.05)
    student = copy.deepcopy(teacher)
    with torch.no_grad():
        for p in student.parameters():
            p.add_(0.02 * torch.randn_like(p))
    teacher.freeze()
    return student, teacher


def _batch(B=2, hw=8, T=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"image": r(B, 4, hw, hw), "prompt_embeds": r(B, T, 12), "negative_prompt_embeds": r(B, T, 12),
            "pooled_prompt_embeds": r(B, 10), "negative_pooled_prompt_embeds": r(B, 10)}


def _disc():
    return nn.Sequential(nn.Conv2d(4, 8, 4, 2, 1, bias=False), nn.SiLU(True), nn.Conv2d(8, 1, 4, 1, 0, bias=False),
                         nn.Flatten())


def _model(gan="lsgan", K=4, dmd=True, seed=0, disc=True):
    student, teacher = _denoisers(seed)
    cfg = FlashDiffusionSD3Config(K=[K], num_iterations_per_K=[100], guidance_scale_min=7.0, guidance_scale_max=13.0,
                                  distill_loss_type="l2", use_dmd_loss=dmd, gan_loss_type=gan,
                                  timestep_distribution="mixture", mixture_num_components=2, mixture_var=0.5)
    mk = lambda cls: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", timestep_spacing="trailing")
    torch.manual_seed(seed + 1)
    return FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                             teacher_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler),
                             sampling_noise_scheduler=mk(FlashFlowMatchEulerDiscreteScheduler),
                             teacher_sampling_noise_scheduler=FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium"),
                             discriminator=_disc() if disc else None)


def _draws(B=2, hw=8, seed=3, start_idx=1):
    g = torch.Generator().manual_seed(seed)
    return {"noise": torch.randn(B, 4, hw, hw, generator=g), "start_idx": start_idx, "guidance": 9.5,
            "dmd_noise": torch.randn(B, 4, hw, hw, generator=g), "dmd_index": torch.tensor([137, 902][:B]),
            "dmd_guidance": 11.0, "gan_noise": torch.randn(B, 4, hw, hw, generator=g),
            "gan_choice": torch.tensor([2, 0][:B])}


def test_flow_match_grids_match_oracle():
    s = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium", timestep_spacing="trailing")
    ts, sig = O3.training_grid()
    assert torch.equal(s.timesteps, ts) and torch.equal(s.sigmas, sig)
    assert s.timesteps[0] == 1000.0 and abs(float(s.sigmas[-1]) - 3 / 1002) < 1e-6       # shift 3 at s = 1/1000
    for K in (1, 4, 32):
        s.set_timesteps(K)
        ts, sig = O3.inference_grid(K)
        assert torch.equal(s.timesteps, ts) and torch.equal(s.sigmas, sig)
        assert len(s.timesteps) == K and s.sigmas[-1] == 0 and s.timesteps[0] == 1000.0
    up = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium")                               # upstream spacing
    up.set_timesteps(4)
    ts, sig = O3.inference_grid(4, spacing="linspace")
    assert torch.equal(up.timesteps, ts) and torch.equal(up.sigmas, sig)


def test_euler_rollout_integrates_a_straight_flow_exactly():
    """For the straight flow x_t = (1-s) x0 + s e the velocity is e - x0 everywhere: Euler from any level lands on x0."""
    s = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium", timestep_spacing="trailing")
    s.set_timesteps(8)
    x0, e = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    x = float(s.sigmas[3]) * e + (1 - float(s.sigmas[3])) * x0
    for t in s.timesteps[3:]:
        assert torch.allclose(s.scale_noise(x0, t, e), x, atol=1e-5)
        x = s.step(e - x0, t, x)[0]
    assert torch.allclose(x, x0, atol=1e-5)
    with pytest.raises(ValueError):
        s.step(e, 123.456, x)


def test_flash_sampler_renoises_to_the_next_level():
    s = FlashFlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium", timestep_spacing="trailing")
    s.set_timesteps(4)
    x0, e = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    t = s.timesteps[1]
    x = s.scale_noise(x0, t, e)
    g = torch.Generator().manual_seed(5)
    prev, pred = s.step(e - x0, t, x, generator=g)
    assert torch.allclose(pred, x0, atol=1e-5)
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    assert torch.allclose(prev, (1 - float(s.sigmas[2])) * x0 + float(s.sigmas[2]) * noise, atol=1e-5)
    last = s.step(e - x0, s.timesteps[-1], s.scale_noise(x0, s.timesteps[-1], e))[0]
    assert torch.allclose(last, x0, atol=1e-5)


@pytest.mark.parametrize("gan", ["lsgan", "hinge", "wgan", "non-saturating", "vanilla"])
@pytest.mark.parametrize("step", [0, 1])
def test_forward_matches_oracle_step(gan, step):
    model = _model(gan=gan)
    batch, draws = _batch(), _draws()
    out = model(batch, step=step, draws=draws)
    cond = {"vector": batch["pooled_prompt_embeds"], "crossattn": batch["prompt_embeds"]}
    unc = {"vector": batch["negative_pooled_prompt_embeds"], "crossattn": batch["negative_prompt_embeds"]}
    call = lambda net: (lambda x, t, c: net(x, t, {"cond": c}))
    ref = O3.flash_forward_sd3(call(model.student_denoiser), call(model.teacher_denoiser), model.discriminator,
                               batch["image"], cond, unc, draws, K=4, step=step, gan_loss_type=gan)
    assert torch.allclose(out["teacher_output"], ref["teacher_output"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["student_output"], ref["student_output"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["loss"][0], ref["loss_G"], rtol=1e-4, atol=1e-6)
    if step % 2 == 0:
        assert out["loss"][1] == 0 and ref["loss_D"] == 0
    else:
        assert torch.allclose(out["loss"][1], ref["loss_D"], rtol=1e-4, atol=1e-6)
    assert out["start_timestep"] == 900.0                      # trailing grid of K=4: 1000, 900, 750, 500 (shift 3)


def test_start_idx_zero_uses_pure_noise_and_no_discriminator_returns_scalar():
    model = _model(disc=False, dmd=False)
    batch, draws = _batch(), _draws(start_idx=0)
    out = model(batch, draws=draws)
    assert torch.equal(out["noisy_sample"], draws["noise"])
    assert out["loss"].dim() == 0 and out["start_timestep"] == 1000.0
    # the student output is the velocity's x0 read-out at sigma = 1
    v = model.student_denoiser(draws["noise"], torch.full((2,), 1000.0),
                               {"cond": {"vector": batch["pooled_prompt_embeds"], "crossattn": batch["prompt_embeds"]}})
    assert torch.allclose(out["student_output"], draws["noise"] - v, atol=1e-5)


def test_get_sigmas_and_missing_embeddings():
    model = _model()
    grid = model.teacher_noise_scheduler_copy
    s = model.get_sigmas(grid, grid.timesteps[torch.tensor([0, 500, 999])])
    assert s.shape == (3, 1, 1, 1) and torch.equal(s.flatten(), grid.sigmas[torch.tensor([0, 500, 999])])
    with pytest.raises(ValueError):
        model.get_sigmas(grid, torch.tensor([123.4567]))
    batch = _batch()
    del batch["negative_prompt_embeds"]
    with pytest.raises(KeyError, match="negative_prompt_embeds"):
        model(batch)


def test_pipeline_encode_prompt_is_called_like_the_reference():
    class Pipe:
        def __init__(self):
            self.calls, self.devices = [], []

        def to(self, device):
            self.devices.append(str(device))
            return self

        def encode_prompt(self, **kw):
            self.calls.append(kw)
            b = _batch()
            return (b["prompt_embeds"], b["negative_prompt_embeds"], b["pooled_prompt_embeds"],
                    b["negative_pooled_prompt_embeds"])
    model = _model()
    model.pipeline, model.cpu_offload = Pipe(), True
    batch = {"image": _batch()["image"], "text": ["a", "b"]}
    out = model(batch, draws=_draws())
    ref = _model()(_batch(), draws=_draws())
    assert torch.allclose(out["loss"][0], ref["loss"][0])
    kw = model.pipeline.calls[0]
    assert kw["prompt"] == kw["prompt_2"] == kw["prompt_3"] == ["a", "b"]
    assert kw["negative_prompt"].startswith("deformed, distorted") and kw["negative_prompt"].endswith("NSFW")
    assert kw["do_classifier_free_guidance"] is True and kw["clip_skip"] is False
    assert model.pipeline.devices == ["cpu", "cpu"]            # moved to the latent's device, then offloaded


def test_undrawn_path_consumes_the_generator_in_reference_order():
    """noise (randn_like z) -> start index (multinomial) -> guidance (rand) -> dmd noise, dmd index (randint, cpu),
    dmd guidance (rand) -> gan noise, gan slot (multinomial): reference :250-286, :427-470, :512-536."""
    model = _model()
    batch = _batch()
    torch.manual_seed(11)
    out = model(batch, step=0)
    torch.manual_seed(11)
    z = batch["image"]
    noise = torch.randn_like(z)
    start_idx = int(torch.multinomial(model._start_index_pmf(4, 0), 1))
    g = float(torch.rand(1)) * 6.0 + 7.0
    dmd_noise = torch.randn_like(z)
    dmd_index = torch.randint(0, 1000, (2,), device="cpu")
    dmd_g = float(torch.rand(1)) * 6.0 + 7.0
    gan_noise = torch.randn_like(z)
    choice = torch.tensor([0.25] * 4).multinomial(2, replacement=True)
    model2 = _model()
    out2 = model2(batch, step=0, draws={"noise": noise, "start_idx": start_idx, "guidance": g, "dmd_noise": dmd_noise,
                                        "dmd_index": dmd_index, "dmd_guidance": dmd_g, "gan_noise": gan_noise,
                                        "gan_choice": choice})
    assert torch.allclose(out["loss"][0], out2["loss"][0], rtol=1e-5)
    assert torch.equal(out["teacher_output"], out2["teacher_output"])


def test_lora_student_trains_only_adapters_with_reference_targets():
    """examples/train_flash_sd3.py:101-120: LoRA on attention, feed-forward AND every AdaLN / embedder linear."""
    model = _model()
    student = copy.deepcopy(model.teacher_denoiser)
    targets = ["to_q", "to_k", "to_v", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2", "proj",
               "linear", "linear_1", "linear_2"]
    names = {n for n, m in student.named_modules() if isinstance(m, nn.Linear)}
    hit = {n for n in names if any(n == t or n.endswith("." + t) for t in targets)}
    assert any(n.endswith("norm1.linear") for n in hit) and any(n.endswith("timestep_embedder.linear_1") for n in hit)
    assert not any("add_q_proj" in n or "to_add_out" in n or "context_embedder" in n for n in hit)
    inject_lora(student, LoraConfig(r=4, lora_alpha=4, target_modules=targets))
    trainable = [n for n, p in student.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)


def test_training_step_update_invariants():
    """reference tests/test_flash/test_flash_diffusion.py:155-187 applied to the SD3 objective: the generator and
    discriminator optimizers move the student and the discriminator; the frozen teacher never moves."""
    model = _model()
    pipe = TrainingPipeline(model, TrainingConfig(optimizers_name=["AdamW", "AdamW"], learning_rates=[1e-3, 1e-3],
                                                  trainable_params=[["student_denoiser"], ["discriminator."]]))
    pipe.configure_optimizers()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    out = pipe.training_step(_batch(), 0)
    assert out["loss_optimizer_0"] > 0 and out["loss_optimizer_1"] > 0
    changed = {n for n, p in model.named_parameters() if not torch.equal(p, before[n])}
    assert any(n.startswith("student_denoiser") for n in changed) and any(n.startswith("discriminator") for n in changed)
    assert not any(n.startswith("teacher_denoiser") for n in changed)


def test_sample_four_steps_and_teacher_reference():
    model = _model()
    z = torch.randn(2, 4, 8, 8)
    out, ref = model.sample(z, num_steps=4, guidance_scale=1.0, teacher_guidance_scale=5.0,
                            conditioner_inputs=_batch(), log_teacher_samples=True,
                            generator=torch.Generator().manual_seed(0))
    assert out.shape == z.shape and ref.shape == z.shape and torch.isfinite(out).all() and torch.isfinite(ref).all()
    assert len(model.sampling_noise_scheduler.timesteps) == 4


def test_inject_lora_wraps_the_patch_convolution_like_peft():
    """The "proj" target of the DiT scripts also names `pos_embed.proj` (a Conv2d): peft gives it a `lora.Conv2d`
    (lora_A with the base kernel / stride, lora_B 1x1).  Any other convolution a target list reaches is refused."""
    from flash.models.lora import LoRAConv2d
    from flash.recipes import DIT_LORA_TARGETS
    student = SD3TransformerOracle(**TINY)
    inject_lora(student, LoraConfig(r=4, lora_alpha=8, target_modules=DIT_LORA_TARGETS))
    proj = student.pos_embed.proj
    assert isinstance(proj, LoRAConv2d) and proj.scaling == 2.0
    sd = student.state_dict()
    assert sd["pos_embed.proj.lora_A.default.weight"].shape == (4, 4, 2, 2)
    assert sd["pos_embed.proj.lora_B.default.weight"].shape == (16, 4, 1, 1)
    assert sd["pos_embed.proj.base_layer.weight"].shape == (16, 4, 2, 2)
    assert proj.lora_A["default"].stride == (2, 2) and float(proj.lora_B["default"].weight.abs().max()) == 0.0
    net = nn.Sequential()
    net.add_module("body", nn.Conv2d(3, 3, 3))
    with pytest.raises(NotImplementedError, match="convolution"):
        inject_lora(net, LoraConfig(r=4, target_modules=["body"]))


def test_sd3_log_samples():
    model = _model()
    logs = model.log_samples(_batch(), input_shape=(4, 8, 8), num_steps=[1, 4], max_samples=8, log_teacher_samples=True)
    assert set(logs) == {"samples_1_steps/FlashFlowMatchEulerDiscreteScheduler_1.0_cfg/student",
                         "samples_4_steps/FlashFlowMatchEulerDiscreteScheduler_1.0_cfg/student",
                         "samples_1_steps/FlowMatchEulerDiscreteScheduler_5.0_cfg/teacher",
                         "samples_4_steps/FlowMatchEulerDiscreteScheduler_5.0_cfg/teacher"}
    assert all(v.shape == (2, 4, 8, 8) and torch.isfinite(v).all() for v in logs.values())
