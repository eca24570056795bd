"""Parity of the rectified-flow (SD3) objective against vectors the REFERENCE ITSELF produced
(tests/golden/reference_sd3_step.pt, written by tests/golden/make_reference_sd3_golden.py from the unmodified
src/flash/models/flash_sd3/flash_diffusion_model.py:187-662 with every random draw recorded):
oracle/flash_step_sd3.py and the product's FlashDiffusionSD3 host logic replay the draws on CPU."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = torch.load(os.path.join(HERE, "golden", "reference_sd3_step.pt"), weights_only=False)
CASES = list(GOLD["cases"])
SCALES = (1.0, 0.7, 0.3)


def _models():
    import make_reference_sd3_golden as G3
    return G3.build_models(GOLD["model_seed"])


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _check_grads(rec, case, student, disc, loss_G, loss_D):
    if case["step"] % 2 == 0:
        loss_G.backward()
        got = {n: p.grad for n, p in student.named_parameters() if p.grad is not None}
        assert set(got) == set(rec["grad_norms"])
        for n, g in rec["grads"].items():
            assert _rel(got[n], g) < 1e-3, (n, _rel(got[n], g))
        for n, gn in rec["grad_norms"].items():
            assert abs(float(got[n].norm()) - float(gn)) <= 1e-3 * float(gn) + 1e-7, n
    else:
        assert torch.allclose(torch.as_tensor(loss_D).detach(), rec["loss_D"], rtol=2e-4, atol=1e-6)
        loss_D.backward()
        for n, p in disc.named_parameters():
            assert _rel(p.grad, rec["grads"]["disc." + n]) < 1e-3, n


@pytest.mark.parametrize("name", CASES)
def test_oracle_sd3_step_matches_reference_run(name):
    from oracle import flash_step_sd3 as O3
    rec = GOLD["cases"][name]
    case, b = rec["case"], GOLD["batch"]
    student, teacher, disc = _models()
    cond = {"vector": b["pooled_prompt_embeds"], "crossattn": b["prompt_embeds"]}
    unc = {"vector": b["negative_pooled_prompt_embeds"], "crossattn": b["negative_prompt_embeds"]}
    call = lambda net: (lambda x, t, c: net(x, t, {"cond": c}))
    out = O3.flash_forward_sd3(call(student), call(teacher), disc, b["image"], cond, unc, rec["draws"], K=GOLD["K"],
                               step=case["step"], gan_loss_type=case["gan"], scales=SCALES,
                               use_teacher_as_real=case["teacher_real"])
    assert _rel(out["student_output"], rec["student_output"]) < 1e-5
    assert _rel(out["teacher_output"], rec["teacher_output"]) < 1e-4
    assert torch.allclose(out["loss_G"].detach(), rec["loss_G"], rtol=2e-4, atol=1e-6), (out["loss_G"], rec["loss_G"])
    _check_grads(rec, case, student, disc, out["loss_G"], out["loss_D"])


@pytest.mark.parametrize("name", CASES)
def test_product_sd3_host_logic_matches_reference_run(name):
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    from flash.schedulers import FlashFlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteScheduler
    rec = GOLD["cases"][name]
    case = rec["case"]
    student, teacher, disc = _models()
    cfg = FlashDiffusionSD3Config(
        K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], guidance_scale_min=7.0, guidance_scale_max=13.0,
        distill_loss_type="l2", timestep_distribution="mixture", mixture_num_components=4, mixture_var=0.5,
        use_dmd_loss=True, dmd_loss_scale=SCALES[1], distill_loss_scale=SCALES[0], adversarial_loss_scale=SCALES[2],
        gan_loss_type=case["gan"], mode_probs=[[0.25, 0.25, 0.25, 0.25]], use_teacher_as_real=case["teacher_real"],
        input_key="image")
    mk = lambda cls, **kw: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", **kw)
    model = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                              teacher_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler, timestep_spacing="trailing"),
                              sampling_noise_scheduler=mk(FlashFlowMatchEulerDiscreteScheduler,
                                                          timestep_spacing="trailing"),
                              discriminator=disc)
    batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in GOLD["batch"].items()}
    out = model(batch, step=case["step"], draws=dict(rec["draws"]))
    assert float(out["start_timestep"]) == rec["start_timestep"]
    assert torch.allclose(out["noisy_sample"], rec["noisy_sample"], rtol=1e-5, atol=1e-6)
    assert _rel(out["student_output"], rec["student_output"]) < 1e-5
    assert _rel(out["teacher_output"], rec["teacher_output"]) < 1e-4
    assert torch.allclose(torch.as_tensor(out["loss"][0]).detach(), rec["loss_G"], rtol=2e-4, atol=1e-6)
    _check_grads(rec, case, student, disc, out["loss"][0], out["loss"][1])


class _ReplayRandn:
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
        assert exc[0] is not None or not self.queue, f"{len(self.queue)} recorded draws were not consumed"


@pytest.mark.parametrize("name", list(GOLD["sample"]))
def test_product_sd3_sampler_matches_reference_run(name):
    """`FlashDiffusionSD3.sample()` (reference flash_sd3/flash_diffusion_model.py:683-843): flash sampling on the
    re-noising flow-matching scheduler, CFG, `max_samples`, the teacher's Euler reference samples."""
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    from flash.schedulers import FlashFlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteScheduler
    rec = GOLD["sample"][name]
    assert rec["other_draws"] == [] and rec["randn_like"] == []
    student, teacher, _ = _models()
    cfg = FlashDiffusionSD3Config(K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], input_key="image")
    mk = lambda cls, **kw: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", **kw)
    model = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                              teacher_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler, timestep_spacing="trailing"),
                              sampling_noise_scheduler=mk(FlashFlowMatchEulerDiscreteScheduler,
                                                          timestep_spacing="trailing"),
                              teacher_sampling_noise_scheduler=mk(FlowMatchEulerDiscreteScheduler),
                              discriminator=None)
    cin = {k: v for k, v in GOLD["batch"].items() if k != "image"}
    with _ReplayRandn(rec["randn"]):
        smp, smp_ref = model.sample(rec["z"].clone(), conditioner_inputs=cin, **rec["kwargs"])
    assert model.sampling_noise_scheduler.timesteps.tolist() == rec["timesteps"].tolist()
    assert smp.shape == rec["sample"].shape and _rel(smp, rec["sample"]) < 1e-4, _rel(smp, rec["sample"])
    if rec["sample_ref"] is None:
        assert smp_ref is None
    else:
        assert _rel(smp_ref, rec["sample_ref"]) < 1e-4


DISTILL = torch.load(os.path.join(HERE, "golden", "reference_sd3_distill.pt"), weights_only=False)


@pytest.mark.parametrize("i", range(len(DISTILL["cases"])))
def test_product_sd3_distill_loss_matches_reference_run(i):
    """`FlashDiffusionSD3._distill_loss` (reference flash_sd3/flash_diffusion_model.py:373-413): l2, l1 and the lpips
    branch with its crop CLAMPED to the latent size (the epsilon model's is not) — same stand-in VAE / perceptual
    distance on both sides (tests/golden/make_reference_sd3_distill_golden.py)."""
    import make_reference_sd3_distill_golden as GD
    import make_reference_step_golden as G
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    from flash.schedulers import FlowMatchEulerDiscreteScheduler
    rec = DISTILL["cases"][i]
    student, teacher, _ = _models()
    cfg = FlashDiffusionSD3Config(K=[GOLD["K"]], num_iterations_per_K=[10 ** 9], input_key="image")
    sched = FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium",
                                                            subfolder="scheduler", timestep_spacing="trailing")
    m = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                          sampling_noise_scheduler=None, discriminator=None)
    m.__dict__["vae"], m.__dict__["lpips"] = G.StubVAE(), G.StubLPIPS()
    m.distill_loss_type = rec["kind"]
    s_out, t_out = GD.latents(rec["shape"], rec["seed"])
    s_out.requires_grad_(True)
    loss = m._distill_loss(s_out, t_out)
    assert torch.allclose(loss, rec["loss"], rtol=1e-5), (rec["kind"], rec["shape"], float(loss), float(rec["loss"]))
    loss.backward()
    assert torch.isfinite(s_out.grad).all() and float(s_out.grad.abs().sum()) > 0
