"""Parity against vectors THE REFERENCE ITSELF produced (tests/golden/reference_step.pt, written by
tests/golden/make_reference_step_golden.py: the unmodified reference `FlashDiffusion.forward`,
src/flash/models/flash/flash_diffusion_model.py:179-667, and `ConditionerWrapper`, imported from /root/reference/src in
the build container, every random draw recorded).

  * CPU: oracle/flash_step.py replays the recorded draws  -> pins the ORACLE's restatement of the step to the reference;
  * CPU: the product's host logic (FlashDiffusion.forward + scheduler classes + ConditionerWrapper, oracle denoisers)
         replays them -> pins the PRODUCT's step logic to the reference directly;
  * GPU: the product's CUDA path (own UNet engine, fused step kernels, CUDA-graph teacher) replays them, bf16 tolerance.

Six cases: lsgan / hinge / vanilla / non-saturating / wgan, generator and discriminator turns, start_idx == 0 (pure
noise, :243-246) and > 0, l1 / l2 distillation, use_teacher_as_real.  Weights are a pure function of a seed
(make_golden.seeded_state_dict) and are regenerated here.
"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

GOLD = torch.load(os.path.join(HERE, "golden", "reference_step.pt"), weights_only=False)
CASES = list(GOLD["cases"])
SCALES = (1.0, 0.7, 0.3)            # distill / dmd / adversarial scales of the generating script


def _models():
    import make_reference_step_golden as G
    return G.build_models(GOLD["model_seed"])


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("name", CASES)
def test_oracle_step_matches_reference_run(name):
    from oracle import flash_step as OF
    rec = GOLD["cases"][name]
    case, draws = rec["case"], rec["draws"]
    student, teacher, disc = _models()
    cond, uncond = {"cond": rec["cond"]}, {"cond": rec["uncond"]}
    out = OF.flash_forward(student, teacher, disc, GOLD["batch"]["image"], cond, uncond, draws, K=GOLD["K"],
                           step=case["step"], use_dmd=case["dmd"], gan_loss_type=case["gan"],
                           distill_type=case["distill"], scales=SCALES, use_teacher_as_real=case["teacher_real"])
    # fp32 on both sides; the CFG combination (w up to 13) amplifies summation-order differences of the two teacher calls
    assert _rel(out["student_output"], rec["student_output"]) < 1e-5
    assert _rel(out["teacher_output"], rec["teacher_output"]) < 1e-4
    assert torch.allclose(out["loss_G"].detach(), rec["loss_G"], rtol=2e-4, atol=1e-6), (out["loss_G"], rec["loss_G"])
    if case["step"] % 2 == 0:
        out["loss_G"].backward()
        got = {n: p.grad for n, p in student.named_parameters() if p.grad is not None}
        assert set(rec["grad_norms"]) == set(got)
        for n, g in rec["grads"].items():
            assert _rel(got[n], g) < 1e-3, (n, _rel(got[n], g))
        for n, gn in rec["grad_norms"].items():
            assert abs(float(got[n].norm()) - float(gn)) <= 1e-3 * float(gn) + 1e-7, n
    else:
        assert torch.allclose(torch.as_tensor(out["loss_D"]).detach(), rec["loss_D"], rtol=2e-4, atol=1e-6)
        out["loss_D"].backward()
        for n, p in disc.named_parameters():
            assert _rel(p.grad, rec["grads"]["disc." + n]) < 1e-3, n
        if case["gan"] == "wgan":       # the reference clamps the critic's weights in place (:573-576)
            for k, v in disc.state_dict().items():
                assert torch.equal(v, rec["disc_state_after_clip"][k]), k


def _product_model(case, student, teacher, disc, device="cpu"):
    from flash.models.embedders import ConditionerWrapper, TorchNNEmbedder, TorchNNEmbedderConfig
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    from flash.schedulers import DPMSolverMultistepScheduler, LCMScheduler
    cfg = FlashDiffusionConfig(
        K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], guidance_scale_min=3.0, guidance_scale_max=13.0,
        distill_loss_type=case["distill"], ucg_keys=["text_emb", "pooled_emb"], timestep_distribution="mixture",
        mixture_num_components=4, mixture_var=0.5, use_dmd_loss=case["dmd"], dmd_loss_scale=SCALES[1],
        distill_loss_scale=SCALES[0], adversarial_loss_scale=SCALES[2], gan_loss_type=case["gan"],
        mode_probs=[[0.25, 0.25, 0.25, 0.25]], use_teacher_as_real=case["teacher_real"], use_empty_prompt=False,
        input_key="image")
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}], ucg_rate=0.0)
    conditioner = ConditionerWrapper([TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
                                      TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident))])
    sched = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0",
                                                        subfolder="scheduler", timestep_spacing="trailing")
    lcm = LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                       timestep_spacing="trailing")
    return FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                          sampling_noise_scheduler=lcm, vae=None, conditioner=conditioner,
                          discriminator=disc).to(device)


@pytest.mark.parametrize("name", CASES)
def test_product_host_logic_matches_reference_run(name):
    """FlashDiffusion.forward of the PRODUCT (2B-batched CFG, scheduler classes, fused-step fallbacks for CPU tensors)
    with the oracle denoisers plugged in, on the reference's recorded draws."""
    rec = GOLD["cases"][name]
    case = rec["case"]
    student, teacher, disc = _models()
    model = _product_model(case, student, teacher, disc)
    cw = model.conditioner(GOLD["batch"], set_ucg_rate_zero=True)["cond"]
    un = model.conditioner(GOLD["batch"], ucg_keys=["text_emb", "pooled_emb"])["cond"]
    for k in rec["cond"]:               # reference ConditionerWrapper outputs (conditioners_wrapper.py:39-90)
        assert torch.equal(cw[k], rec["cond"][k]) and torch.equal(un[k], rec["uncond"][k]), k
    out = model({k: v.clone() for k, v in GOLD["batch"].items()}, step=case["step"], draws=dict(rec["draws"]))
    assert out["start_timestep"] == rec["start_timestep"]
    assert torch.allclose(out["noisy_sample"], rec["noisy_sample"], rtol=1e-5, atol=1e-6)
    assert _rel(out["student_output"], rec["student_output"]) < 1e-5
    assert _rel(out["teacher_output"], rec["teacher_output"]) < 1e-4
    assert torch.allclose(torch.as_tensor(out["loss"][0]).detach(), rec["loss_G"], rtol=2e-4, atol=1e-6)
    if case["step"] % 2 == 0:
        assert out["loss"][1] == 0
        out["loss"][0].backward()
        got = {n: p.grad for n, p in student.named_parameters() if p.grad is not None}
        for n, g in rec["grads"].items():
            assert _rel(got[n], g) < 1e-3, (n, _rel(got[n], g))
    else:
        assert torch.allclose(torch.as_tensor(out["loss"][1]).detach(), rec["loss_D"], rtol=2e-4, atol=1e-6)
        out["loss"][1].backward()
        for n, p in disc.named_parameters():
            assert _rel(p.grad, rec["grads"]["disc." + n]) < 1e-3, n


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_step_matches_reference_run(name):
    """The CUDA path (product UNet engine in bf16, fused step kernels, CUDA-graph teacher rollout) on the reference's
    recorded draws.  Tolerances are bf16's: the reference ran the same weights in fp32."""
    from flash.models.lora import LoraConfig
    from flash.models.unets import DiffusersUNet2DCondWrapper
    rec = GOLD["cases"][name]
    case = rec["case"]
    o_student, o_teacher, o_disc = _models()
    dev = torch.device("cuda", 0)
    teacher = DiffusersUNet2DCondWrapper(**GOLD["unet_kwargs"])
    teacher.load_state_dict(o_teacher.state_dict())
    student = DiffusersUNet2DCondWrapper(**GOLD["unet_kwargs"])
    student.load_state_dict(o_teacher.state_dict())
    student.add_adapter(LoraConfig(**GOLD["lora"]))
    student.load_state_dict(o_student.state_dict())
    teacher, student, disc = teacher.to(dev), student.to(dev), o_disc.to(dev)
    teacher.freeze()
    model = _product_model(case, student, teacher, disc, device=dev)
    batch = {k: v.to(dev) for k, v in GOLD["batch"].items()}
    draws = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in rec["draws"].items()}
    out = model(batch, step=case["step"], draws=draws)
    assert out["start_timestep"] == rec["start_timestep"]
    assert _rel(out["student_output"].cpu(), rec["student_output"]) < 3e-2
    assert _rel(out["teacher_output"].cpu(), rec["teacher_output"]) < 3e-2
    lg, ref_g = float(out["loss"][0]), float(rec["loss_G"])
    assert abs(lg - ref_g) <= 5e-2 * abs(ref_g) + 1e-3, (lg, ref_g)
    if case["step"] % 2 == 0:
        out["loss"][0].backward()
        got = {n: p.grad for n, p in student.named_parameters() if p.grad is not None}
        cos = []
        for n, g in rec["grads"].items():
            a, b = got[n].float().cpu().reshape(-1), g.reshape(-1)
            cos.append(float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)))
        cos = torch.tensor(cos)
        assert cos.median() > 0.99 and cos.min() > 0.9, (cos.median(), cos.min())
    else:
        ld, ref_d = float(out["loss"][1]), float(rec["loss_D"])
        assert abs(ld - ref_d) <= 5e-2 * abs(ref_d) + 1e-3, (ld, ref_d)


class _ReplayRandn:
    """torch.randn replaced by the tensors the reference run drew, in order (moved to the requested device)."""

    def __init__(self, tensors):
        self.queue, self.orig = list(tensors), torch.randn

    def __enter__(self):
        def randn(*shape, **k):
            t = self.queue.pop(0)
            want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(t.shape) == want, (t.shape, want)
            return t.to(device=k.get("device") or "cpu", dtype=k.get("dtype") or t.dtype)
        torch.randn = randn
        return self

    def __exit__(self, *exc):
        torch.randn = self.orig
        assert not exc[0] is None or not self.queue, f"{len(self.queue)} recorded draws were not consumed"


def _sampler_model(student, teacher, device="cpu"):
    from flash.models.embedders import ConditionerWrapper, TorchNNEmbedder, TorchNNEmbedderConfig
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    from flash.schedulers import DPMSolverMultistepScheduler, LCMScheduler
    cfg = FlashDiffusionConfig(K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], ucg_keys=["text_emb", "pooled_emb"],
                               input_key="image")
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}], ucg_rate=0.0)
    conditioner = ConditionerWrapper([TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
                                      TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident))])
    mk = lambda cls: cls.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                         timestep_spacing="trailing")
    return FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher,
                          teacher_noise_scheduler=mk(DPMSolverMultistepScheduler),
                          teacher_sampling_noise_scheduler=mk(DPMSolverMultistepScheduler),
                          sampling_noise_scheduler=mk(LCMScheduler), vae=None, conditioner=conditioner,
                          discriminator=None).to(device)


@pytest.mark.parametrize("name", list(GOLD["sample"]))
def test_product_sampler_matches_reference_run(name):
    """`sample()` (reference :754-915): few-step LCM sampling with CFG, `max_samples`, the teacher reference samples."""
    rec = GOLD["sample"][name]
    student, teacher, _ = _models()
    model = _sampler_model(student, teacher)
    cin = {k: v for k, v in GOLD["batch"].items() if k != "image"}
    with _ReplayRandn(rec["randn"]):
        smp, smp_ref = model.sample(rec["z"].clone(), conditioner_inputs=dict(cin), **rec["kwargs"])
    assert model.sampling_noise_scheduler.timesteps.tolist() == rec["lcm_timesteps"].tolist()
    assert smp.shape == rec["sample"].shape and _rel(smp, rec["sample"]) < 1e-4, _rel(smp, rec["sample"])
    if rec["sample_ref"] is None:
        assert smp_ref is None
    else:
        assert _rel(smp_ref, rec["sample_ref"]) < 1e-4


def test_product_log_samples_matches_reference_run():
    """`log_samples()` (reference :917-1019): same keys, same tensors, same order of latent draws."""
    rec = GOLD["log_samples"]
    student, teacher, _ = _models()
    model = _sampler_model(student, teacher)
    cin = {k: v for k, v in GOLD["batch"].items() if k != "image"}
    with _ReplayRandn(rec["randn"]):
        logs = model.log_samples(dict(cin), input_shape=(4, 16, 16), guidance_scale=1.0, teacher_guidance_scale=3.0,
                                 max_samples=8, num_steps=[1, 2], device="cpu", log_teacher_samples=True)
    assert set(logs) == set(rec["logs"])
    for k, v in rec["logs"].items():
        assert logs[k].shape == v.shape and _rel(logs[k], v) < 1e-4, (k, _rel(logs[k], v))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GOLD["sample"]))
def test_cuda_sampler_matches_reference_run(name):
    from flash.models.lora import LoraConfig
    from flash.models.unets import DiffusersUNet2DCondWrapper
    rec = GOLD["sample"][name]
    o_student, o_teacher, _ = _models()
    dev = torch.device("cuda", 0)
    teacher = DiffusersUNet2DCondWrapper(**GOLD["unet_kwargs"])
    teacher.load_state_dict(o_teacher.state_dict())
    student = DiffusersUNet2DCondWrapper(**GOLD["unet_kwargs"])
    student.load_state_dict(o_teacher.state_dict())
    student.add_adapter(LoraConfig(**GOLD["lora"]))
    student.load_state_dict(o_student.state_dict())
    teacher.freeze(); student.freeze()
    model = _sampler_model(student.to(dev), teacher.to(dev), device=dev)
    cin = {k: v.to(dev) for k, v in GOLD["batch"].items() if k != "image"}
    with _ReplayRandn(rec["randn"]):
        smp, smp_ref = model.sample(rec["z"].to(dev), conditioner_inputs=cin, **rec["kwargs"])
    assert _rel(smp.cpu(), rec["sample"]) < 4e-2, _rel(smp.cpu(), rec["sample"])
    if rec["sample_ref"] is not None:
        assert _rel(smp_ref.cpu(), rec["sample_ref"]) < 4e-2


def test_product_training_pipeline_matches_reference_run():
    """`TrainingPipeline.configure_optimizers` + two `training_step`s (reference src/flash/trainer/trainer.py:76-218:
    regex parameter groups, per optimizer a full forward with step=i and fresh draws, zero_grad / backward / step, SGD
    with momentum on the discriminator): same frozen set, same parameter updates."""
    from flash.trainer import TrainingConfig, TrainingPipeline
    rec = GOLD["trainer"]
    student, teacher, disc = _models()
    case = dict(distill="l2", dmd=True, gan="lsgan", teacher_real=False)
    model = _product_model(case, student, teacher, disc)
    pipe = TrainingPipeline(model, TrainingConfig(**rec["config"]))
    pipe.configure_optimizers()
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == rec["trainable"]
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    for it, st in enumerate(rec["steps"]):
        out = pipe.training_step({k: v.clone() for k, v in GOLD["batch"].items()}, it, draws=[dict(d) for d in st["draws"]])
        assert out["start_timestep"] == st["start_timestep"]
        for k in ("loss_optimizer_0", "loss_optimizer_1"):
            assert torch.allclose(torch.as_tensor(out[k]).float(), st[k], rtol=1e-3, atol=1e-6), (it, k, out[k], st[k])
    after = {n: p.detach() for n, p in model.named_parameters()}
    moved = sorted(n for n in before if not torch.equal(before[n], after[n]))
    assert moved == rec["moved"]
    for n, d in rec["deltas"].items():
        assert _rel(after[n] - before[n], d) < 2e-3, (n, _rel(after[n] - before[n], d))
    for n, dn in rec["delta_norms"].items():
        got = float((after[n] - before[n]).norm())
        assert abs(got - float(dn)) <= 2e-3 * float(dn) + 1e-9, n


def test_product_start_index_pmf_matches_reference_run():
    """`_get_timesteps` (reference :139-177): the start-index distribution (uniform / gaussian / mixture with mode
    probabilities) the reference handed to torch.multinomial, and the timestep it looked up."""
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    from flash.schedulers import DPMSolverMultistepScheduler
    student, teacher, _ = _models()
    for rec in GOLD["pmf"]:
        cfg = FlashDiffusionConfig(K=[rec["K"]], num_iterations_per_K=[10 ** 9], timestep_distribution=rec["dist"],
                                   ucg_keys=["text_emb"], input_key="image", **rec["kw"])
        sched = DPMSolverMultistepScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0",
                                                            subfolder="scheduler", timestep_spacing="trailing")
        m = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                           sampling_noise_scheduler=None, vae=None, conditioner=None, discriminator=None)
        pmf = m._start_index_pmf(rec["K"], 0)
        assert torch.allclose(pmf, rec["prob"], rtol=1e-6, atol=1e-9), (rec["dist"], rec["K"], rec["kw"])
        idx, t0 = m._get_timesteps(num_samples=3, K=rec["K"], K_step=0, start_idx=rec["start_idx"])
        assert torch.equal(t0, rec["start_timestep"]) and sched.timesteps.tolist() == rec["timesteps"].tolist()


def test_product_lpips_distill_glue_matches_reference_run():
    """`_distill_loss(..., "lpips")` (reference :383-397): centre crop of 64x64 latents, decode, clamp, distance, mean —
    the same stand-in VAE / perceptual distance on both sides, so only the reference's glue is compared."""
    import make_reference_step_golden as G
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    rec = GOLD["lpips_glue"]
    student, teacher, _ = _models()
    cfg = FlashDiffusionConfig(K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], ucg_keys=["text_emb"], input_key="image")
    m = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=None,
                       sampling_noise_scheduler=None, vae=None, conditioner=None, discriminator=None)
    m.__dict__["vae"] = G.StubVAE()
    m.distill_loss_type = "lpips"
    m.__dict__["lpips"] = G.StubLPIPS()
    g = torch.Generator().manual_seed(rec["seed"])
    s_out, t_out = torch.randn(*rec["shape"], generator=g), torch.randn(*rec["shape"], generator=g)
    assert torch.allclose(m._distill_loss(s_out, t_out), rec["loss"], rtol=1e-5)
