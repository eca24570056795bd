"""Golden vectors produced BY THE REFERENCE ITSELF: python tests/golden/make_reference_step_golden.py

Imports the unmodified reference package from /root/reference/src (build container only — it does not travel to the GPU
box) and runs its own `FlashDiffusion.forward` (src/flash/models/flash/flash_diffusion_model.py:179-366, with
`_get_timesteps` :139-177, `_distill_loss` :368-399, `_dmd_loss` :401-499, `_gan_loss` :501-667), its own
`ConditionerWrapper` + `TorchNNEmbedder` (src/flash/models/embedders/conditioners_wrapper.py:39-90) and its own
`sample()` (:754-915) on CPU in fp32, then writes inputs, random draws and outputs to tests/golden/reference_step.pt.
tests/test_reference_golden.py replays them through oracle/flash_step.py (pinning the oracle's restatement of the step),
through the product's host logic on CPU and, on the GPU box, through the CUDA path.

What is and is not the reference here:
  * the step logic, the conditioner wrapper, the sampler loop: REFERENCE code, unmodified;
  * the denoisers: oracle/unet.py modules (the reference's wrappers subclass diffusers models, which cannot be
    installed offline) — the reference only ever calls them through the wrapper signature
    `(sample, timestep, conditioning, down_intrablock_additional_residuals, return_intermediate)`;
  * `diffusers.schedulers`: diffusers is absent, so the module is served by flash-diffusion_b200/flash/schedulers.py
    (loaded by file path).  The scheduler ARITHMETIC therefore stays unpinned upstream; what this fixture adds is that
    reference step code x product scheduler classes == the oracle's independent function-style restatement.
  * every random draw the reference makes is recorded (torch.randn_like / rand / randint / multinomial are wrapped,
    not replaced; `start_idx` alone can be forced so that both the start_idx == 0 and > 0 branches are covered).
"""
import importlib.util
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"

UNET_KWARGS = dict(in_channels=4, out_channels=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"],
                   up_block_types=["CrossAttnUpBlock2D", "UpBlock2D"], block_out_channels=[64, 128],
                   layers_per_block=1, cross_attention_dim=96, transformer_layers_per_block=[1, 2],
                   attention_head_dim=[1, 2], use_linear_projection=True, class_embed_type="projection",
                   projection_class_embeddings_input_dim=48)
LORA = dict(r=8, lora_alpha=8, init_lora_weights="gaussian", target_modules=["to_k", "to_q", "to_v", "to_out.0"])
K = 4
B, HW, T = 2, 16, 7


def install_shims():
    """Stand-ins for the packages the reference imports at module scope and that cannot be installed here."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Absent(torch.nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("placeholder for a diffusers / lpips class that is not installed in this container")

    spec = importlib.util.spec_from_file_location(
        "diffusers.schedulers", os.path.join(ROOT, "flash-diffusion_b200", "flash", "schedulers.py"))
    sched = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sched)
    sys.modules["diffusers.schedulers"] = sched
    d = mod("diffusers", schedulers=sched, T2IAdapter=_Absent, DiffusionPipeline=_Absent)
    d.__path__ = []
    m = mod("diffusers.models", UNet2DConditionModel=_Absent, UNet2DModel=_Absent, AutoencoderKL=_Absent)
    m.__path__ = []
    mod("diffusers.models.transformers", SD3Transformer2DModel=_Absent, Transformer2DModel=_Absent)
    mod("diffusers.models.embeddings", Timesteps=_Absent, TimestepEmbedding=_Absent)
    mod("lpips", LPIPS=_Absent)

    class LightningModule(torch.nn.Module):
        """The five LightningModule services the reference's TrainingPipeline.training_step uses (trainer.py:169-218),
        with Lightning's documented semantics: toggle_optimizer freezes every parameter that belongs to ANOTHER
        optimizer only, untoggle_optimizer restores the flags, manual_backward is loss.backward()."""
        automatic_optimization = True
        global_rank = 0

        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

        def optimizers(self):
            return self.optims

        def toggle_optimizer(self, optimizer):
            mine = {id(p) for g in optimizer.param_groups for p in g["params"]}
            self._toggled = {}
            for opt in self.optims:
                for g in opt.param_groups:
                    for p in g["params"]:
                        if id(p) not in mine and id(p) not in self._toggled:
                            self._toggled[id(p)] = (p, p.requires_grad)
                            p.requires_grad = False

        def untoggle_optimizer(self, optimizer):
            for p, flag in self._toggled.values():
                p.requires_grad = flag
            self._toggled = {}

        def manual_backward(self, loss):
            loss.backward()

    pl = mod("pytorch_lightning", LightningModule=LightningModule, Trainer=_Absent)
    pl.__path__ = []
    mod("pytorch_lightning.callbacks", Callback=object)
    mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    return sched


class Tape:
    """Records every random tensor the reference draws, in order; `force_start_idx` overrides torch.multinomial(prob, 1)
    (the start index of `_get_timesteps`) only."""

    def __init__(self, force_start_idx=None):
        self.events, self.force, self.probs = [], force_start_idx, []
        self.orig = dict(randn_like=torch.randn_like, rand=torch.rand, randint=torch.randint,
                         multinomial=torch.multinomial, t_multinomial=torch.Tensor.multinomial)

    def __enter__(self):
        o, ev, tape = self.orig, self.events, self

        def randn_like(*a, **k):
            r = o["randn_like"](*a, **k); ev.append(("randn_like", r.clone())); return r

        def rand(*a, **k):
            r = o["rand"](*a, **k); ev.append(("rand", r.clone())); return r

        def randint(*a, **k):
            r = o["randint"](*a, **k); ev.append(("randint", r.clone())); return r

        def multinomial(p, n, *a, **k):
            r = o["multinomial"](p, n, *a, **k)
            tape.probs.append(p.clone())
            if tape.force is not None and n == 1 and p.dim() == 1:
                r = torch.tensor([tape.force])
            ev.append(("multinomial", r.clone())); return r

        def t_multinomial(self_, n, *a, **k):
            r = o["t_multinomial"](self_, n, *a, **k); ev.append(("tensor.multinomial", r.clone())); return r

        torch.randn_like, torch.rand, torch.randint, torch.multinomial = randn_like, rand, randint, multinomial
        torch.Tensor.multinomial = t_multinomial
        return self

    def __exit__(self, *exc):
        o = self.orig
        torch.randn_like, torch.rand, torch.randint, torch.multinomial = o["randn_like"], o["rand"], o["randint"], o["multinomial"]
        torch.Tensor.multinomial = o["t_multinomial"]

    def draws(self, g_min, g_max):
        """The recorded draws in the oracle's vocabulary; the two guidance scales are the reference's affine map of its
        uniform draws (:284-286, :453-457)."""
        kinds = [k for k, _ in self.events]
        want = ["randn_like", "multinomial", "rand", "randn_like", "randint", "rand", "randn_like", "tensor.multinomial"]
        assert kinds == want, kinds            # the order flash_diffusion_model.py:236,167,285,416,418,454,515,527 draws in
        v = [t for _, t in self.events]
        return dict(noise=v[0], start_idx=int(v[1]), guidance=float(v[2] * (g_max - g_min) + g_min), dmd_noise=v[3],
                    dmd_timestep=v[4], dmd_guidance=float(v[5] * (g_max - g_min) + g_min), gan_noise=v[6],
                    guidance_uniform=float(v[2]), dmd_guidance_uniform=float(v[5]),
                    gan_timesteps=torch.tensor([10, 250, 500, 750])[v[7]])


class StubLPIPS(torch.nn.Module):
    """stands where lpips.LPIPS(net="vgg") stands in `_distill_loss` (:383-397): any perceptual distance [B,1,1,1]"""

    def forward(self, a, b):
        return ((a - b) ** 2).mean((1, 2, 3), keepdim=True) + (a - b).abs().amax((1, 2, 3), keepdim=True)


class StubVAE:
    """stands where AutoencoderKLDiffusers stands in `_distill_loss`: latents [B,4,h,w] -> images [B,3,2h,2w]"""

    def decode(self, z):
        return torch.nn.functional.interpolate(z[:, :3] * 1.5 + 0.25 * z[:, 3:4], scale_factor=2, mode="nearest")


class RandnTape:
    """records torch.randn (the LCM scheduler's inter-step noise and log_samples' initial latents)"""

    def __init__(self):
        self.events, self.orig = [], torch.randn

    def __enter__(self):
        def randn(*a, **k):
            r = self.orig(*a, **k); self.events.append(r.clone()); return r
        torch.randn = randn
        return self

    def __exit__(self, *exc):
        torch.randn = self.orig


def build_models(seed):
    """Weights are a pure function of the seed (tests/golden/make_golden.py::seeded_state_dict): regenerated by the
    tests, not stored."""
    from make_golden import seeded_state_dict
    from oracle.unet import LoraConfig, UNet2DConditionOracle
    teacher = UNet2DConditionOracle(**UNET_KWARGS)
    teacher.load_state_dict(seeded_state_dict(teacher, seed))
    student = UNet2DConditionOracle(**UNET_KWARGS)
    student.load_state_dict(teacher.state_dict())
    student.add_adapter(LoraConfig(**LORA))
    sd = seeded_state_dict(student, seed + 1)
    sd.update({k: v for k, v in teacher.state_dict().items() if k in sd and "lora" not in k})
    # peft keeps the wrapped layer under `base_layer`: those entries are the teacher's weights
    for k in list(sd):
        if ".base_layer." in k:
            sd[k] = teacher.state_dict()[k.replace(".base_layer.", ".")]
    student.load_state_dict(sd)
    teacher.freeze()
    disc = build_disc(seed + 2)
    return student, teacher, disc


def build_disc(seed):
    from make_golden import seeded_state_dict
    disc = torch.nn.Sequential(torch.nn.Conv2d(128, 8, 4, 2, 1, bias=False), torch.nn.SiLU(),
                               torch.nn.Conv2d(8, 1, 4, 1, 0, bias=False), torch.nn.Flatten())
    disc.load_state_dict(seeded_state_dict(disc, seed))
    return disc


def main():
    assert os.path.isdir(REF_SRC), "the reference tree is only present in the build container"
    sched_mod = install_shims()
    sys.path.insert(0, ROOT)            # oracle/
    sys.path.insert(0, HERE)            # make_golden.seeded_state_dict
    sys.path.insert(0, REF_SRC)         # the reference's `flash` package (NOT flash-diffusion_b200/flash)
    from flash.models.embedders import ConditionerWrapper, TorchNNEmbedder, TorchNNEmbedderConfig
    from flash.models.flash import FlashDiffusion, FlashDiffusionConfig
    import flash
    assert flash.__path__[0].startswith(REF_SRC) or os.path.realpath(flash.__path__[0]).startswith(REF_SRC)

    cases = [
        dict(name="lsgan_dmd_G", gan="lsgan", step=0, start_idx=2, dmd=True, distill="l2", teacher_real=False),
        dict(name="lsgan_dmd_D", gan="lsgan", step=1, start_idx=1, dmd=True, distill="l2", teacher_real=False),
        dict(name="start0_hinge_G", gan="hinge", step=0, start_idx=0, dmd=True, distill="l1", teacher_real=False),
        dict(name="vanilla_D_teacher_real", gan="vanilla", step=1, start_idx=3, dmd=True, distill="l2", teacher_real=True),
        dict(name="nonsat_G_free_start", gan="non-saturating", step=0, start_idx=None, dmd=True, distill="l2",
             teacher_real=False),
        dict(name="wgan_D", gan="wgan", step=1, start_idx=2, dmd=True, distill="l2", teacher_real=False),
    ]
    out = dict(unet_kwargs=UNET_KWARGS, lora=LORA, K=K, cases={}, generated_by=os.path.relpath(__file__, ROOT),
               reference_files=["src/flash/models/flash/flash_diffusion_model.py",
                                "src/flash/models/embedders/conditioners_wrapper.py",
                                "src/flash/models/embedders/torch_nn/embedders.py"])
    out["model_seed"] = 4242
    student, teacher, disc = build_models(out["model_seed"])
    student_state = {k: v.clone() for k, v in student.state_dict().items()}
    disc_state = {k: v.clone() for k, v in disc.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    batch = dict(image=torch.randn(B, 4, HW, HW, generator=g), text_emb=torch.randn(B, T, 96, generator=g),
                 pooled_emb=torch.randn(B, 48, generator=g))
    out["batch"] = {k: v.clone() for k, v in batch.items()}

    for ci, case in enumerate(cases):
        student.load_state_dict(student_state); disc.load_state_dict(disc_state)
        for p in list(student.parameters()) + list(disc.parameters()):
            p.grad = None
        cfg = FlashDiffusionConfig(
            K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=3.0, guidance_scale_max=13.0,
            distill_loss_type=case["distill"], ucg_keys=["text_emb", "pooled_emb"], timestep_distribution="mixture",
            mixture_num_components=4, mixture_var=0.5, use_dmd_loss=case["dmd"], dmd_loss_scale=0.7,
            distill_loss_scale=1.0, adversarial_loss_scale=0.3, gan_loss_type=case["gan"],
            mode_probs=[[0.25, 0.25, 0.25, 0.25]], use_teacher_as_real=case["teacher_real"], use_empty_prompt=False,
            input_key="image")
        conditioner = ConditionerWrapper([
            TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", nn_modules=["torch.nn.Identity"],
                                                  nn_modules_kwargs=[{}], ucg_rate=0.0)),
            TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", nn_modules=["torch.nn.Identity"],
                                                  nn_modules_kwargs=[{}], ucg_rate=0.0))])
        tsched = sched_mod.DPMSolverMultistepScheduler.from_pretrained(
            "stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler", timestep_spacing="trailing")
        lcm = sched_mod.LCMScheduler.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                                     timestep_spacing="trailing")
        model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=tsched,
                               sampling_noise_scheduler=lcm, vae=None, conditioner=conditioner, discriminator=disc)
        model.switch_teacher = False           # attribute the fork reads at :228 but never sets
        tsched.set_timesteps(K)
        torch.manual_seed(100 + ci)
        with Tape(case["start_idx"]) as tape:
            res = model(dict(batch), step=case["step"])
        draws = tape.draws(3.0, 13.0)
        loss_G, loss_D = res["loss"]
        rec = dict(case=case, draws=draws, loss_G=torch.as_tensor(float(loss_G)), loss_D=torch.as_tensor(float(loss_D)),
                   student_output=res["student_output"].detach().clone(),
                   teacher_output=res["teacher_output"].detach().clone(),
                   noisy_sample=res["noisy_sample"].detach().clone(), start_timestep=int(res["start_timestep"]))
        if case["step"] % 2 == 0:
            loss_G.backward()
            named = [(n, p) for n, p in student.named_parameters() if p.grad is not None]
            rec["grad_norms"] = {n: p.grad.norm().clone() for n, p in named}
            rec["grads"] = {n: p.grad.clone() for i, (n, p) in enumerate(named) if i % 4 == 0}   # every 4th tensor in full
        else:
            loss_D.backward()
            rec["grads"] = {"disc." + n: p.grad.clone() for n, p in disc.named_parameters() if p.grad is not None}
        if case["gan"] == "wgan":
            rec["disc_state_after_clip"] = {k: v.clone() for k, v in disc.state_dict().items()}
        # the conditioner outputs the reference wrapper produced (cond / ucg-forced zeros)
        rec["cond"] = {k: v.clone() for k, v in conditioner(batch, set_ucg_rate_zero=True)["cond"].items()}
        rec["uncond"] = {k: v.clone() for k, v in conditioner(batch, ucg_keys=cfg.ucg_keys)["cond"].items()}
        out["cases"][case["name"]] = rec
        print(case["name"], "start_idx", draws["start_idx"], "t0", rec["start_timestep"], "loss_G", float(loss_G),
              "loss_D", float(loss_D), "grads", len(rec["grads"]))

    # ---- `_get_timesteps` (:139-177): the start-index pmf of every distribution type, K = 32 and 4
    out["pmf"] = []
    for dist, kw in [("uniform", {}), ("gaussian", {}),
                     ("mixture", dict(mixture_num_components=4, mixture_var=0.5, mode_probs=[[0.25, 0.25, 0.25, 0.25]])),
                     ("mixture", dict(mixture_num_components=4, mixture_var=2.0, mode_probs=[[0.4, 0.2, 0.2, 0.2]])),
                     ("mixture", dict(mixture_num_components=2, mixture_var=0.5, mode_probs=[[0.0, 1.0]]))]:
        for K_ in (32, 4):
            cfg = FlashDiffusionConfig(K=[K_], num_iterations_per_K=[10 ** 9], timestep_distribution=dist,
                                       ucg_keys=["text_emb"], input_key="image", **kw)
            tsched = sched_mod.DPMSolverMultistepScheduler.from_pretrained(
                "stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler", timestep_spacing="trailing")
            m = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=tsched,
                               sampling_noise_scheduler=None, vae=None, conditioner=None, discriminator=None)
            torch.manual_seed(5)
            with Tape() as tape:
                idx, t0 = m._get_timesteps(num_samples=3, K=K_, K_step=0, device="cpu")
            out["pmf"].append(dict(dist=dist, K=K_, kw=kw, prob=tape.probs[0].clone(), start_idx=int(idx),
                                   start_timestep=t0.clone(), timesteps=tsched.timesteps.clone()))
    print("pmf cases", len(out["pmf"]))

    # ---- `_distill_loss` lpips branch (:383-397): center crop 64x64, decode both, clamp, perceptual distance, mean
    cfg = FlashDiffusionConfig(K=[K], num_iterations_per_K=[10 ** 9], ucg_keys=["text_emb"], input_key="image")
    m = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=None,
                       sampling_noise_scheduler=None, vae=StubVAE(), conditioner=None, discriminator=None)
    m.distill_loss_type, m.lpips = "lpips", StubLPIPS()
    gl = torch.Generator().manual_seed(3)
    s_out, t_out = torch.randn(2, 4, 72, 80, generator=gl), torch.randn(2, 4, 72, 80, generator=gl)
    out["lpips_glue"] = dict(seed=3, shape=(2, 4, 72, 80), loss=m._distill_loss(s_out.clone(), t_out.clone()).clone())
    print("lpips glue", float(out["lpips_glue"]["loss"]))

    # ---- the reference's TrainingPipeline.configure_optimizers / training_step (src/flash/trainer/trainer.py:76-218):
    #      two optimizers, per optimizer a full forward with step=i and fresh draws, zero_grad / backward / step
    from flash.trainer.trainer import TrainingPipeline
    from flash.trainer.training_config import TrainingConfig
    student.load_state_dict(student_state); disc.load_state_dict(disc_state)
    for p in student.parameters():
        p.requires_grad_(True)                  # configure_optimizers must do the freezing by regex itself
    for n, p in student.named_parameters():
        if "lora" not in n:
            p.requires_grad_(False)             # (what add_adapter leaves: only LoRA weights trainable)
    cfg = FlashDiffusionConfig(
        K=[K], num_iterations_per_K=[10 ** 9], guidance_scale_min=3.0, guidance_scale_max=13.0,
        distill_loss_type="l2", ucg_keys=["text_emb", "pooled_emb"], timestep_distribution="mixture",
        mixture_num_components=4, mixture_var=0.5, use_dmd_loss=True, dmd_loss_scale=0.7, distill_loss_scale=1.0,
        adversarial_loss_scale=0.3, gan_loss_type="lsgan", mode_probs=[[0.25, 0.25, 0.25, 0.25]], input_key="image")
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}], ucg_rate=0.0)
    conditioner = ConditionerWrapper([TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
                                      TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident))])
    mk = lambda cls: cls.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                         timestep_spacing="trailing")
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher,
                           teacher_noise_scheduler=mk(sched_mod.DPMSolverMultistepScheduler),
                           sampling_noise_scheduler=mk(sched_mod.LCMScheduler), vae=None, conditioner=conditioner,
                           discriminator=disc)
    model.switch_teacher = False
    tcfg = dict(optimizers_name=["SGD", "SGD"], learning_rates=[2e-3, 5e-3], optimizers_kwargs=[{}, {"momentum": 0.5}],
                trainable_params=[["student_denoiser"], ["discriminator."]], lr_schedulers_name=[None, None],
                lr_schedulers_kwargs=[{}, {}], lr_schedulers_interval=["step", "step"], lr_schedulers_frequency=[1, 1])
    pipe = TrainingPipeline(model, TrainingConfig(**tcfg))
    pipe.configure_optimizers()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    torch.manual_seed(321)
    steps = []
    for it, forced in enumerate([(2, 1), (3, 0)]):          # two training steps: momentum and the turn order matter
        tapes = []

        class Both:                                          # one tape per model.forward call (two per training_step)
            def __enter__(self_):
                return self_
        orig_forward = model.forward

        def taped_forward(*a, _orig=orig_forward, **k):
            with Tape(forced[len(tapes)]) as tp:
                r = _orig(*a, **k)
            tapes.append(tp)
            return r
        model.forward = taped_forward
        outs = pipe.training_step(dict(batch), it)
        model.forward = orig_forward
        steps.append(dict(draws=[tp.draws(3.0, 13.0) for tp in tapes],
                          loss_optimizer_0=torch.as_tensor(float(outs["loss_optimizer_0"])),
                          loss_optimizer_1=torch.as_tensor(float(outs["loss_optimizer_1"])),
                          start_timestep=int(outs["start_timestep"])))
        print("training_step", it, {k: float(v) for k, v in outs.items() if k.startswith("loss")})
    after = {n: p.detach().clone() for n, p in model.named_parameters()}
    moved = [n for n in before if not torch.equal(before[n], after[n])]
    assert set(moved) <= set(trainable), "a frozen parameter moved"
    out["trainer"] = dict(config=tcfg, steps=steps, trainable=trainable, moved=sorted(moved),
                          delta_norms={n: (after[n] - before[n]).norm().clone() for n in moved},
                          deltas={n: (after[n] - before[n]).clone() for i, n in enumerate(sorted(moved)) if i % 4 == 0})
    print("trainer: trainable", len(trainable), "moved", len(moved))

    # ---- the reference's sampler (:754-915) and sample logging (:917-1019)
    student.load_state_dict(student_state)
    cfg = FlashDiffusionConfig(K=[K], num_iterations_per_K=[10 ** 9], ucg_keys=["text_emb", "pooled_emb"], input_key="image")
    ident = dict(nn_modules=["torch.nn.Identity"], nn_modules_kwargs=[{}], ucg_rate=0.0)
    conditioner = ConditionerWrapper([TorchNNEmbedder(TorchNNEmbedderConfig(input_key="text_emb", **ident)),
                                      TorchNNEmbedder(TorchNNEmbedderConfig(input_key="pooled_emb", **ident))])
    mk = lambda cls: cls.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0", subfolder="scheduler",
                                         timestep_spacing="trailing")
    model = FlashDiffusion(cfg, student_denoiser=student, teacher_denoiser=teacher,
                           teacher_noise_scheduler=mk(sched_mod.DPMSolverMultistepScheduler),
                           teacher_sampling_noise_scheduler=mk(sched_mod.DPMSolverMultistepScheduler),
                           sampling_noise_scheduler=mk(sched_mod.LCMScheduler), vae=None, conditioner=conditioner,
                           discriminator=None)
    cin = {k: v for k, v in batch.items() if k != "image"}
    z = torch.randn(B, 4, HW, HW, generator=g)
    out["sample"] = {}
    for name, kw in [("lcm4_cfg1", dict(num_steps=4, guidance_scale=1.0)),
                     ("lcm2_cfg1.7_teacher", dict(num_steps=2, guidance_scale=1.7, teacher_guidance_scale=5.0,
                                                  log_teacher_samples=True)),
                     ("lcm1_max1", dict(num_steps=1, guidance_scale=1.0, max_samples=1))]:
        torch.manual_seed(55)
        with RandnTape() as tape:
            smp, smp_ref = model.sample(z.clone(), conditioner_inputs=dict(cin), **kw)
        out["sample"][name] = dict(kwargs=kw, z=z.clone(), randn=tape.events, sample=smp.clone(),
                                   sample_ref=None if smp_ref is None else smp_ref.clone(),
                                   lcm_timesteps=model.sampling_noise_scheduler.timesteps.clone())
        print("sample", name, "noise draws", len(tape.events), "timesteps", model.sampling_noise_scheduler.timesteps.tolist())
    torch.manual_seed(56)
    with RandnTape() as tape:
        logs = model.log_samples(dict(cin), input_shape=(4, HW, HW), guidance_scale=1.0, teacher_guidance_scale=3.0,
                                 max_samples=8, num_steps=[1, 2], device="cpu", log_teacher_samples=True)
    out["log_samples"] = dict(randn=tape.events, logs={k: v.clone() for k, v in logs.items()})
    print("log_samples keys", sorted(logs))

    path = os.path.join(HERE, "reference_step.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
