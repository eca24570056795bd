"""Golden vectors produced by the REFERENCE's own `FlashDiffusionSD3.forward`
(src/flash/models/flash_sd3/flash_diffusion_model.py:187-371, `_dmd_loss` :415-496, `_gan_loss` :498-662, `get_sigmas`
:947-958), imported unmodified from /root/reference/src in the build container:
    python tests/golden/make_reference_sd3_golden.py   ->   tests/golden/reference_sd3_step.pt

As in make_reference_step_golden.py: the denoisers are oracle/sd3.py MMDiTs, `diffusers.schedulers` is served by
flash-diffusion_b200/flash/schedulers.py (so the run also shows that the product's flow-matching scheduler classes offer
every attribute the reference touches: deepcopy of the untouched scheduler as the 1000-level training grid,
`.sigmas` / `.timesteps` look-ups in get_sigmas, `.step` inside the rollout), the `pipeline.encode_prompt` of the
reference (a diffusers StableDiffusion3Pipeline) is a stand-in that returns the batch's precomputed embeddings, and every
random draw is recorded.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_reference_step_golden as G  # noqa: E402

SD3_KWARGS = dict(sample_size=8, patch_size=2, in_channels=4, num_layers=2, attention_head_dim=8, num_attention_heads=2,
                  joint_attention_dim=12, caption_projection_dim=16, pooled_projection_dim=10, out_channels=4,
                  pos_embed_max_size=8)
K, B, HW, T = 4, 2, 8, 5


def build_models(seed):
    from make_golden import seeded_state_dict
    from oracle.sd3 import SD3TransformerOracle

    def seeded(net, s):
        sd = seeded_state_dict(net, s)
        for name, buf in net.named_buffers():          # deterministic position tables stay as constructed
            if name in sd:
                sd[name] = buf.clone()
        net.load_state_dict(sd)
        return net
    teacher = seeded(SD3TransformerOracle(**SD3_KWARGS), seed)
    student = seeded(SD3TransformerOracle(**SD3_KWARGS), seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in student.parameters():                  # a full-rank perturbation of the teacher (no LoRA here)
            p.add_(0.02 * torch.randn(p.shape, generator=g))
    teacher.freeze()
    disc = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 4, 2, 1, bias=False), torch.nn.SiLU(),
                               torch.nn.Conv2d(8, 1, 4, 1, 0, bias=False), torch.nn.Flatten())
    disc.load_state_dict(seeded_state_dict(disc, seed + 2))
    return student, teacher, disc


class Pipeline:
    """stand-in for the diffusers SD3 pipeline the reference calls at :197-222"""

    def __init__(self, batch):
        self.batch = batch

    def to(self, *a, **k):
        return self

    def encode_prompt(self, **kw):
        b = self.batch
        return (b["prompt_embeds"], b["negative_prompt_embeds"], b["pooled_prompt_embeds"],
                b["negative_pooled_prompt_embeds"])


def draws_sd3(tape, g_min, g_max):
    kinds = [k for k, _ in tape.events]
    assert kinds == ["randn_like", "multinomial", "rand", "randn_like", "randint", "rand", "randn_like",
                     "tensor.multinomial"], kinds
    v = [t for _, t in tape.events]
    return dict(noise=v[0], start_idx=int(v[1]), guidance=float(v[2] * (g_max - g_min) + g_min), dmd_noise=v[3],
                dmd_index=v[4], dmd_guidance=float(v[5] * (g_max - g_min) + g_min), gan_noise=v[6], gan_choice=v[7])


def main():
    sched_mod = G.install_shims()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, G.REF_SRC)
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)

    g = torch.Generator().manual_seed(11)
    r = lambda *s: torch.randn(*s, generator=g)
    batch = {"image": r(B, 4, HW, HW), "prompt_embeds": r(B, T, 12), "negative_prompt_embeds": r(B, T, 12),
             "pooled_prompt_embeds": r(B, 10), "negative_pooled_prompt_embeds": r(B, 10), "text": ["a", "b"]}
    out = dict(sd3_kwargs=SD3_KWARGS, K=K, model_seed=777, cases={}, generated_by=os.path.relpath(__file__, ROOT),
               batch={k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()},
               reference_files=["src/flash/models/flash_sd3/flash_diffusion_model.py"])
    student, teacher, disc = build_models(out["model_seed"])
    s_state = {k: v.clone() for k, v in student.state_dict().items()}
    d_state = {k: v.clone() for k, v in disc.state_dict().items()}
    cases = [dict(name="lsgan_G", gan="lsgan", step=0, start_idx=1, teacher_real=False),
             dict(name="lsgan_D", gan="lsgan", step=1, start_idx=2, teacher_real=False),
             dict(name="start0_hinge_G", gan="hinge", step=0, start_idx=0, teacher_real=False),
             dict(name="vanilla_D_teacher_real", gan="vanilla", step=1, start_idx=3, teacher_real=True),
             dict(name="wgan_G_free_start", gan="wgan", step=0, start_idx=None, teacher_real=False)]
    for ci, case in enumerate(cases):
        student.load_state_dict(s_state); disc.load_state_dict(d_state)
        for p in list(student.parameters()) + list(disc.parameters()):
            p.grad = None
        cfg = FlashDiffusionSD3Config(
            K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=7.0, guidance_scale_max=13.0,
            distill_loss_type="l2", timestep_distribution="mixture", mixture_num_components=4, mixture_var=0.5,
            use_dmd_loss=True, dmd_loss_scale=0.7, distill_loss_scale=1.0, adversarial_loss_scale=0.3,
            gan_loss_type=case["gan"], mode_probs=[[0.25, 0.25, 0.25, 0.25]], use_teacher_as_real=case["teacher_real"],
            input_key="image")
        mk = lambda cls, **kw: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", **kw)
        model = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                                  teacher_noise_scheduler=mk(sched_mod.FlowMatchEulerDiscreteScheduler,
                                                             timestep_spacing="trailing"),
                                  sampling_noise_scheduler=mk(sched_mod.FlashFlowMatchEulerDiscreteScheduler,
                                                              timestep_spacing="trailing"),
                                  vae=None, conditioner=None, discriminator=disc, pipeline=Pipeline(batch))
        model.switch_teacher = False
        torch.manual_seed(200 + ci)
        with G.Tape(case["start_idx"]) as tape:
            res = model(dict(batch), step=case["step"])
        draws = draws_sd3(tape, 7.0, 13.0)
        loss_G, loss_D = res["loss"]
        rec = dict(case=case, draws=draws, loss_G=torch.as_tensor(float(loss_G)), loss_D=torch.as_tensor(float(loss_D)),
                   student_output=res["student_output"].detach().clone(),
                   teacher_output=res["teacher_output"].detach().clone(),
                   noisy_sample=res["noisy_sample"].detach().clone(), start_timestep=float(res["start_timestep"]))
        if case["step"] % 2 == 0:
            loss_G.backward()
            named = [(n, p) for n, p in student.named_parameters() if p.grad is not None]
            rec["grad_norms"] = {n: p.grad.norm().clone() for n, p in named}
            rec["grads"] = {n: p.grad.clone() for i, (n, p) in enumerate(named) if i % 6 == 0}
        else:
            loss_D.backward()
            rec["grads"] = {"disc." + n: p.grad.clone() for n, p in disc.named_parameters() if p.grad is not None}
        out["cases"][case["name"]] = rec
        print(case["name"], "start_idx", draws["start_idx"], "t0", rec["start_timestep"], "loss_G", float(loss_G),
              "loss_D", float(loss_D))
    # ---- the reference's SD3 sampler (:683-843): few-step flash sampling (re-noising Euler), CFG, teacher reference samples
    student.load_state_dict(s_state)
    cfg = FlashDiffusionSD3Config(K=[K], num_iterations_per_K=[10 ** 9], input_key="image")
    mk = lambda cls, **kw: cls.from_pretrained("stabilityai/stable-diffusion-3-medium", subfolder="scheduler", **kw)
    model = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher,
                              teacher_noise_scheduler=mk(sched_mod.FlowMatchEulerDiscreteScheduler,
                                                         timestep_spacing="trailing"),
                              sampling_noise_scheduler=mk(sched_mod.FlashFlowMatchEulerDiscreteScheduler,
                                                          timestep_spacing="trailing"),
                              teacher_sampling_noise_scheduler=mk(sched_mod.FlowMatchEulerDiscreteScheduler),
                              vae=None, conditioner=None, discriminator=None, pipeline=Pipeline(batch))
    z0 = torch.randn(B, 4, HW, HW, generator=g)
    out["sample"] = {}
    for name, kw in [("flash4_cfg1", dict(num_steps=4, guidance_scale=1.0)),
                     ("flash2_cfg2.5_teacher", dict(num_steps=2, guidance_scale=2.5, teacher_guidance_scale=5.0,
                                                    log_teacher_samples=True)),
                     ("flash1_max1", dict(num_steps=1, guidance_scale=1.0, max_samples=1))]:
        torch.manual_seed(77)
        with G.RandnTape() as tape, G.Tape() as tape2:
            smp, smp_ref = model.sample(z0.clone(), conditioner_inputs={"text": batch["text"]}, **kw)
        out["sample"][name] = dict(kwargs=kw, z=z0.clone(), randn=tape.events,
                                   randn_like=[t for k, t in tape2.events if k == "randn_like"],
                                   other_draws=[k for k, _ in tape2.events if k != "randn_like"],
                                   sample=smp.clone(), sample_ref=None if smp_ref is None else smp_ref.clone(),
                                   timesteps=model.sampling_noise_scheduler.timesteps.clone())
        print("sd3 sample", name, "randn", len(tape.events), "randn_like", len(out["sample"][name]["randn_like"]),
              model.sampling_noise_scheduler.timesteps.tolist())

    path = os.path.join(HERE, "reference_sd3_step.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
