"""Golden values of the REFERENCE's own `FlashDiffusionSD3._distill_loss` (src/flash/models/flash_sd3/
flash_diffusion_model.py:373-413) for its three branches — l2, l1 and lpips (centre crop of AT MOST 64x64 latents,
clamped to the latent size; decode both; clamp to [-1, 1]; perceptual distance; mean) — run unmodified from
/root/reference/src in the build container:
    python tests/golden/make_reference_sd3_distill_golden.py   ->   tests/golden/reference_sd3_distill.pt

The VAE and the perceptual distance are the stand-ins of make_reference_step_golden.py on both sides (the glue is what is
pinned here; the VAE / LPIPS arithmetic has its own oracles).  Latent shapes cover: larger than the crop in both axes,
smaller in one, smaller in both."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_reference_step_golden as G  # noqa: E402
import make_reference_sd3_golden as S  # noqa: E402

SHAPES = [(2, 4, 72, 80), (2, 4, 40, 72), (1, 4, 16, 24)]


def latents(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)


def main():
    sched_mod = G.install_shims()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, G.REF_SRC)
    from flash.models.flash_sd3 import FlashDiffusionSD3, FlashDiffusionSD3Config
    import flash
    assert os.path.realpath(flash.__path__[0]).startswith(G.REF_SRC)
    student, teacher, _ = S.build_models(777)
    sched = sched_mod.FlowMatchEulerDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-3-medium",
                                                                      subfolder="scheduler", timestep_spacing="trailing")
    cfg = FlashDiffusionSD3Config(K=[S.K], num_iterations_per_K=[10 ** 9], input_key="image")
    m = FlashDiffusionSD3(cfg, student_denoiser=student, teacher_denoiser=teacher, teacher_noise_scheduler=sched,
                          sampling_noise_scheduler=None, vae=G.StubVAE(), conditioner=None, discriminator=None)
    m.lpips = G.StubLPIPS()
    out = dict(cases=[], generated_by=os.path.relpath(__file__, ROOT),
               reference_files=["src/flash/models/flash_sd3/flash_diffusion_model.py:373-413"])
    for si, shape in enumerate(SHAPES):
        for kind in ("l2", "l1", "lpips"):
            m.distill_loss_type = kind
            s_out, t_out = latents(shape, 40 + si)
            loss = m._distill_loss(s_out.clone(), t_out.clone())
            out["cases"].append(dict(shape=shape, seed=40 + si, kind=kind, loss=loss.detach().clone()))
            print(shape, kind, float(loss))
    torch.save(out, os.path.join(HERE, "reference_sd3_distill.pt"))


if __name__ == "__main__":
    main()
