"""ORACLE (test infrastructure) — the Flash-Diffusion objective restated as one plain function.

Follows reference src/flash/models/flash/flash_diffusion_model.py line by line:
  :236-257 noising, :260-280 student prediction + x0, :284-324 teacher CFG rollout (two separate B-sized teacher
  calls, DPM-Solver++), :328 c_skip/c_out mix, :368-399 distill loss, :401-499 DMD loss, :501-667 GAN loss.
Every random draw is an explicit input (SURVEY.md §8c decision 5).
PINNED: tests/test_reference_golden.py replays tests/golden/reference_step.pt — six runs of the reference's own
`FlashDiffusion.forward` (imported from /root/reference/src by tests/golden/make_reference_step_golden.py, draws recorded)
— through this function: outputs, both losses and the LoRA / discriminator gradients agree to fp32 rounding.  (The
denoiser and scheduler arithmetic inside stays unpinned upstream, see oracle/unet.py.)
"""
import torch
import torch.nn.functional as F

from . import schedulers as S


def scalings(t, sigma_data=0.5):
    s = t / 0.1
    return sigma_data ** 2 / (s ** 2 + sigma_data ** 2), s / (s ** 2 + sigma_data ** 2) ** 0.5


def predicted_x0(eps, t, x_t, ac):
    a = torch.as_tensor(ac, dtype=x_t.dtype, device=x_t.device)[t.long()].view(-1, 1, 1, 1)
    return (x_t - (1 - a).sqrt() * eps) / a.sqrt()


def flash_forward(student, teacher, discriminator, z, cond, uncond, draws, *, K=32, step=0, use_dmd=True,
                  gan_loss_type="lsgan", distill_type="l2", scales=(1.0, 1.0, 1.0), use_teacher_as_real=False,
                  add_noise_mode="closed_form", distill_fn=None):
    """Returns dict(loss_G, loss_D, student_output, teacher_output, distill, dmd, gan_G).
    draws: noise, start_idx, guidance, dmd_noise, dmd_timestep, dmd_guidance, gan_noise, gan_timesteps."""
    ac = S.alphas_cumprod()
    ts = S.trailing_timesteps(K)

    def add_noise(x, eps, t):        # off-schedule DMD / GAN timesteps: both upstream readings (decision (1), §8c)
        if add_noise_mode == "schedule_index":
            return S.add_noise_schedule_index(ac, x, eps, t, K)
        return S.add_noise(ac, x, eps, t)
    B = z.shape[0]
    start_idx = int(draws["start_idx"])
    t0 = torch.full((B,), int(ts[start_idx]), device=z.device, dtype=torch.long)
    noise = draws["noise"]
    x_t = noise if start_idx == 0 else S.add_noise(ac, z, noise, t0)
    eps_s = student(x_t, t0.float(), cond)
    c_skip, c_out = scalings(t0.float())
    c_skip, c_out = c_skip.view(-1, 1, 1, 1), c_out.view(-1, 1, 1, 1)
    x0_s = predicted_x0(eps_s, t0, x_t, ac)
    w = float(draws["guidance"])

    def eps_fn(x, t):
        tt = torch.full((B,), float(t), device=x.device)
        return w * teacher(x, tt, cond) + (1 - w) * teacher(x, tt, uncond)

    with torch.no_grad():
        teacher_output = S.dpm_rollout(eps_fn, x_t.detach().clone(), ac, K, start_idx)
    student_output = c_skip * x_t + c_out * x0_s
    diff = student_output - teacher_output
    if distill_fn is not None:          # e.g. the lpips branch (oracle/lpips.py), reference :383-397
        distill = distill_fn(student_output, teacher_output)
    else:
        distill = (diff ** 2 if distill_type == "l2" else diff.abs()).reshape(B, -1).mean(1).mean()
    loss = distill * scales[0]
    dmd = torch.zeros((), device=z.device)
    if use_dmd:
        td = draws["dmd_timestep"].long()
        noisy_s = add_noise(student_output, draws["dmd_noise"], td)
        with torch.no_grad():
            wd = float(draws["dmd_guidance"])
            real = wd * teacher(noisy_s, td.float(), cond) + (1 - wd) * teacher(noisy_s, td.float(), uncond)
            fake = student(noisy_s, td.float(), cond)
        a = torch.as_tensor(ac, dtype=z.dtype, device=z.device)[td].view(-1, 1, 1, 1)
        coeff = (real - fake) * (1 - a).sqrt() / a.sqrt()          # (score_fake - score_real), score = -eps
        x0_real = predicted_x0(real, td, noisy_s, ac)
        weight = 1.0 / ((student_output - x0_real).abs().mean([1, 2, 3], keepdim=True) + 1e-5).detach()
        dmd = F.mse_loss(student_output, (student_output - weight * coeff).detach())
        loss = loss + dmd * scales[1]
    tg = draws["gan_timesteps"].long()
    real_x = teacher_output if use_teacher_as_real else z
    fake_x = student_output if step % 2 == 0 else student_output.detach()
    noisy = torch.cat([add_noise(fake_x, draws["gan_noise"], tg), add_noise(real_x, draws["gan_noise"], tg)])
    cond2 = {"cond": {k: torch.cat([v, v]) for k, v in cond["cond"].items()}}
    feats = teacher(noisy, torch.cat([tg, tg]).float(), cond2, return_intermediate=True)
    f_fake, f_real = feats.chunk(2)
    valid, zeros = torch.ones(B, 1, device=z.device), torch.zeros(B, 1, device=z.device)
    loss_G, loss_D = 0, 0
    if gan_loss_type == "wgan":        # reference :573-576: the critic's weights are clipped in place, on BOTH turns
        with torch.no_grad():
            for p in discriminator.parameters():
                p.clamp_(-0.01, 0.01)
    if step % 2 == 0:
        d = discriminator(f_fake)
        loss_G = {"lsgan": lambda: F.mse_loss(torch.sigmoid(d), valid), "hinge": lambda: -d.mean(),
                  "wgan": lambda: -d.mean(),
                  "non-saturating": lambda: -torch.mean(torch.log(torch.sigmoid(d) + 1e-8)),
                  "vanilla": lambda: F.binary_cross_entropy_with_logits(d, valid)}[gan_loss_type]()
    else:
        dr, df = discriminator(f_real), discriminator(f_fake.detach())
        loss_D = {"lsgan": lambda: 0.5 * (F.mse_loss(torch.sigmoid(dr), valid) + F.mse_loss(torch.sigmoid(df), zeros)),
                  "hinge": lambda: F.relu(1.0 - dr).mean() + F.relu(1.0 + df).mean(),
                  "wgan": lambda: -dr.mean() + df.mean(),
                  "non-saturating": lambda: -torch.mean(torch.log(torch.sigmoid(dr) + 1e-8) + torch.log(1 - torch.sigmoid(df) + 1e-8)),
                  "vanilla": lambda: F.binary_cross_entropy_with_logits(dr, valid) + F.binary_cross_entropy_with_logits(df, zeros)}[gan_loss_type]()
    total = loss + scales[2] * loss_G
    return dict(loss_G=total, loss_D=loss_D, student_output=student_output, teacher_output=teacher_output,
                distill=distill, dmd=dmd, gan_G=loss_G)
