"""ORACLE (test infrastructure) — the rectified-flow (SD3) Flash-Diffusion objective restated as one plain function.

Follows reference src/flash/models/flash_sd3/flash_diffusion_model.py line by line: :253-283 noising on the K-point
teacher grid, :288-320 teacher CFG rollout (two separate B-sized teacher calls, Euler x += (sigma'-sigma) v), :322-324
student x0 read-out, :372-383 distill loss, :413-494 DMD loss, :496-658 GAN loss (lsgan / hinge / wgan /
non-saturating / vanilla), :1043-1060 sigma look-up.  Every random draw is an explicit input.

The flow-matching grid is restated from upstream diffusers' published `FlowMatchEulerDiscreteScheduler` (v0.29); the
"trailing" spacing of the authors' fork is read as in flash/schedulers.py (assumption).
PINNED (step logic): tests/test_reference_sd3_golden.py replays tests/golden/reference_sd3_step.pt — five runs of the
reference's own `FlashDiffusionSD3.forward`, imported from /root/reference/src by
tests/golden/make_reference_sd3_golden.py with its draws recorded — through this function: outputs, losses and
gradients agree to fp32 rounding.  Unpinned upstream: the scheduler grid itself and the MMDiT (diffusers is not
installable here), and what the fork's `return_post_mid_blocks=True` backbone call returns (:563; read as the full
backbone output).
"""
import numpy as np
import torch
import torch.nn.functional as F


def shift_map(s, shift):
    return shift * s / (1 + (shift - 1) * s)


def training_grid(n=1000, shift=3.0):
    """(timesteps[n], sigmas[n]) of the scheduler at construction."""
    s = shift_map(torch.from_numpy(np.linspace(1, n, n, dtype=np.float32)[::-1].copy()) / n, shift)
    return s * n, s


def inference_grid(K, n=1000, shift=3.0, spacing="trailing"):
    """(timesteps[K], sigmas[K+1]) after set_timesteps(K)."""
    _, train_sig = training_grid(n, shift)
    if spacing == "trailing":
        raw = np.arange(n, 0, -n / K, dtype=np.float64)[:K]
        s = shift_map(torch.from_numpy((raw / n).astype(np.float32)), shift)
    else:
        raw = np.linspace(float(train_sig[0]) * n, float(train_sig[-1]) * n, K, dtype=np.float32)
        s = shift_map(torch.from_numpy(raw / n), shift)
    return s * n, torch.cat([s, torch.zeros(1)])


def gan_objective(D, f_fake, f_real, B, step, kind):
    valid = torch.ones(B, 1, device=f_fake.device)
    zeros = torch.zeros(B, 1, device=f_fake.device)
    if kind == "wgan":                 # reference flash_sd3/flash_diffusion_model.py:568-571: critic weights clipped in place
        with torch.no_grad():
            for p in D.parameters():
                p.clamp_(-0.01, 0.01)
    if step % 2 == 0:
        d = D(f_fake)
        if kind in ("wgan", "hinge"):
            return -d.mean(), 0
        if kind == "lsgan":
            return F.mse_loss(torch.sigmoid(d), valid), 0
        if kind == "non-saturating":
            return -torch.mean(torch.log(torch.sigmoid(d) + 1e-8)), 0
        return F.binary_cross_entropy_with_logits(d, valid), 0
    d_r, d_f = D(f_real), D(f_fake.detach())
    if kind == "wgan":
        return 0, -d_r.mean() + d_f.mean()
    if kind == "lsgan":
        return 0, 0.5 * (F.mse_loss(torch.sigmoid(d_r), valid) + F.mse_loss(torch.sigmoid(d_f), zeros))
    if kind == "hinge":
        return 0, F.relu(1.0 - d_r).mean() + F.relu(1.0 + d_f).mean()
    if kind == "non-saturating":
        return 0, -torch.mean(torch.log(torch.sigmoid(d_r) + 1e-8) + torch.log(1 - torch.sigmoid(d_f) + 1e-8))
    return 0, F.binary_cross_entropy_with_logits(d_r, valid) + F.binary_cross_entropy_with_logits(d_f, zeros)


def flash_forward_sd3(student, teacher, discriminator, z, cond, uncond, draws, *, K=32, step=0, use_dmd=True,
                      gan_loss_type="lsgan", scales=(1.0, 1.0, 1.0), use_teacher_as_real=False, shift=3.0,
                      spacing="trailing", n_train=1000):
    """student/teacher: callables (x, t[B] float, cond) -> velocity.  draws: noise, start_idx, guidance, dmd_noise,
    dmd_index, dmd_guidance, gan_noise, gan_choice.  Returns dict(loss_G, loss_D, student_output, teacher_output,
    distill, dmd, gan_G)."""
    B = z.shape[0]
    ts, sig = inference_grid(K, n_train, shift, spacing)
    train_ts, train_sig = training_grid(n_train, shift)
    i0 = int(draws["start_idx"])
    noise = draws["noise"]
    s0 = float(sig[i0])
    x_t = noise if i0 == 0 else s0 * noise + (1.0 - s0) * z
    w = float(draws["guidance"])
    x = x_t.detach().clone()
    with torch.no_grad():
        for i in range(i0, K):
            t = torch.full((B,), float(ts[i]), device=z.device)
            v = w * teacher(x, t, cond) + (1 - w) * teacher(x, t, uncond)
            x = x + (float(sig[i + 1]) - float(sig[i])) * v
    teacher_output = x
    t0 = torch.full((B,), float(ts[i0]), device=z.device)
    student_output = x_t - student(x_t, t0, cond) * s0
    distill = ((student_output - teacher_output) ** 2).reshape(B, -1).mean(1).mean()
    loss = distill * scales[0]
    dmd = torch.zeros((), device=z.device)
    if use_dmd:
        idx = torch.as_tensor(draws["dmd_index"]).long()
        td = train_ts[idx].to(z.device)
        sd = train_sig[idx].to(z.device).view(-1, 1, 1, 1)
        noisy_s = sd * draws["dmd_noise"] + (1.0 - sd) * student_output
        with torch.no_grad():
            wd = float(draws["dmd_guidance"])
            real = wd * teacher(noisy_s, td, cond) + (1 - wd) * teacher(noisy_s, td, uncond)
            fake = student(noisy_s, td, cond)
        coeff = (-fake) - (-real)
        weight = 1.0 / ((student_output - real).abs().mean([1, 2, 3], keepdim=True) + 1e-5).detach()
        dmd = F.mse_loss(student_output, (student_output - weight * coeff).detach(), reduction="mean")
        loss = loss + dmd * scales[1]
    gan_G, loss_D = torch.zeros((), device=z.device), 0
    if discriminator is not None:
        slots = [-10, -250, -500, -750]
        pick = [slots[int(c)] for c in torch.as_tensor(draws["gan_choice"]).reshape(-1)]
        tg = torch.stack([train_ts[p] for p in pick]).to(z.device)
        sg = torch.stack([train_sig[p] for p in pick]).to(z.device).view(-1, 1, 1, 1)
        real_img = teacher_output if use_teacher_as_real else z
        noisy = torch.cat([sg * draws["gan_noise"] + (1.0 - sg) * student_output,
                           sg * draws["gan_noise"] + (1.0 - sg) * real_img], dim=0)
        cond2 = {k: torch.cat([v, v], dim=0) for k, v in cond.items()}
        feats = teacher(noisy, torch.cat([tg, tg]), cond2)
        f_fake, f_real = feats.chunk(2, dim=0)
        gan_G, loss_D = gan_objective(discriminator, f_fake, f_real, B, step, gan_loss_type)
        loss = loss + scales[2] * gan_G
    return {"loss_G": loss, "loss_D": loss_D, "student_output": student_output, "teacher_output": teacher_output,
            "distill": distill, "dmd": dmd, "gan_G": gan_G}
