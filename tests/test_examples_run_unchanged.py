"""BASELINE north_star: "keeping the repo's src/flash trainer, embedder and denoiser wrapper API so examples/train_flash_*.py
run unchanged against it".  The UNMODIFIED reference script examples/train_flash_sd.py (BASELINE config 1: SD1.5 UNet
512x512, random-init teacher/student) is executed with `runpy` against this repository's `flash` package plus the
compat shims for the packages that cannot be installed offline (flash-diffusion_b200/compat: diffusers, peft,
pytorch_lightning, braceexpand, lpips).  Its yaml (read by the script from ./configs/flash_sd.yaml) is the reference's
own file with the dataset path pointed at a synthetic webdataset shard and the rollout shortened.

Without a GPU (this container) the run is plumbing-only — everything the script builds (VAE, CLIP conditioner, teacher /
LoRA student, discriminator, FlashDiffusion with the lpips loss, data module with its filter / mapper chain, trainer,
callbacks) plus one batch through the data pipeline (FLASH_MAX_STEPS=0; the denoisers are CUDA-only).  The reference
tree does not travel to the GPU box, so there the test is skipped; the 2-step GPU run of the same script is recorded in
profiles/r02_example_train_flash_sd_gpu.txt."""
import io
import json
import os
import runpy
import sys
import tarfile

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# FLASH_REF_SCRIPT / FLASH_REF_YAML: where a one-off GPU run (tools/example_run_cmd.py) drops the reference's two files
SCRIPT = os.environ.get("FLASH_REF_SCRIPT", "/root/reference/examples/train_flash_sd.py")
REF_YAML = os.environ.get("FLASH_REF_YAML", "/root/reference/examples/configs/flash_sd.yaml")


def make_shard(path, n=6, px=1024):
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(0)
    with tarfile.open(path, "w") as tf:
        for i in range(n):
            img = Image.fromarray(rng.integers(0, 255, (px, px, 3), dtype=np.uint8))
            buf = io.BytesIO()
            img.save(buf, format="JPEG", quality=60)
            meta = json.dumps({"caption": f"synthetic image {i}", "aesthetic_score": 6.5 if i != 1 else 3.0}).encode()
            for name, data in ((f"{i:06d}.jpg", buf.getvalue()), (f"{i:06d}.json", meta)):
                info = tarfile.TarInfo(name)
                info.size = len(data)
                tf.addfile(info, io.BytesIO(data))


def write_config(workdir, shard, steps):
    with open(REF_YAML) as f:
        cfg = yaml.safe_load(f)
    cfg["SHARDS_PATH_OR_URLS"] = [f"pipe:cat {shard}"]
    cfg["K"] = [4, 4, 4, 4]                 # shorter teacher rollout; every other setting (lpips, DMD, lsgan, ...) as shipped
    cfg["BATCH_SIZE"] = 2
    cfg["NUM_STEPS"] = [1]
    cfg["CKPT_EVERY_N_STEPS"] = max(steps, 1)
    os.makedirs(os.path.join(workdir, "configs"), exist_ok=True)
    with open(os.path.join(workdir, "configs", "flash_sd.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    return cfg


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="the reference tree is not on this box")
def test_train_flash_sd_runs_unchanged(tmp_path, monkeypatch):
    import torch
    steps = 2 if torch.cuda.is_available() else 0
    shard = str(tmp_path / "000000.tar")
    make_shard(shard)
    cfg = write_config(str(tmp_path), shard, steps)
    assert cfg["DISTILL_LOSS_TYPE"] == "lpips"
    monkeypatch.chdir(tmp_path)
    for k, v in dict(SLURM_NPROCS="1", SLURM_NNODES="1", FLASH_MAX_STEPS=str(steps)).items():
        monkeypatch.setenv(k, v)
    compat = os.path.join(ROOT, "flash-diffusion_b200", "compat")
    monkeypatch.setattr(sys, "path", sys.path + [compat])          # AFTER site-packages: a real install would win
    runpy.run_path(SCRIPT, run_name="__main__")
    runs = sorted(os.listdir(tmp_path / "logs"))
    assert len(runs) == 1 and runs[0].endswith("-FlashSD15")
    summary = json.load(open(tmp_path / "logs" / runs[0] / "fit_summary.json"))
    assert summary["steps"] == steps
    if steps == 0:
        sb = summary["sanity_batch"]
        assert sb["image"] == [2, 3, 512, 512] and sb["text"].startswith("list")       # mappers: crop, resize, rename
    else:
        assert len(summary["losses"]) == steps and all(l["loss_optimizer_0"] > 0 for l in summary["losses"])
        assert any(f.endswith("_lora.safetensors") for f in os.listdir(tmp_path / "logs" / runs[0] / "checkpoints"))
