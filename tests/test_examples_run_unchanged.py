"""BASELINE north_star: "keeping the repo's src/flash trainer, embedder and denoiser wrapper API so examples/train_flash_*.py
run unchanged against it".  The UNMODIFIED reference script examples/train_flash_sd.py (BASELINE config 1: SD1.5 UNet
512x512, random-init teacher/student) is executed with `runpy` against this repository's `flash` package plus the
compat shims for the packages that cannot be installed offline (flash-diffusion_b200/compat: diffusers, peft,
pytorch_lightning, braceexpand, lpips).  Its yaml (read by the script from ./configs/flash_sd.yaml) is the reference's
own file with the dataset path pointed at a synthetic webdataset shard and the rollout shortened.

The other three training scripts of the BASELINE configs — examples/train_flash_sdxl.py (config 2), train_flash_pixart.py
(config 3), train_flash_sd3.py (config 4) — run unmodified the same way (`test_full_size_scripts_run_unchanged`).  They
build 2-5 billion-parameter teachers / students / text encoders, so on a box without a GPU the script is executed under
`torch.device("meta")`: every constructor, every `load_state_dict` of the HF-layout checkpoint keys (strict=True in the
SD3 script), the per-key weight surgery of the SDXL / PixArt scripts, LoRA injection through `peft.get_peft_model`, the
schedulers' `from_pretrained`, the discriminators, FlashDiffusion(SD3) with the lpips objective, the data module and the
trainer run for real on shapes, without touching the weights' memory; one batch is then pulled through the data pipeline.
On a B200 they run 2 real training steps (`tools/example_run_cmd.py <name>`; profiles/r02_example_train_flash_*_gpu.txt).

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


def write_config(workdir, shard, steps, ref_yaml=None, name="flash_sd.yaml"):
    with open(ref_yaml or REF_YAML) as f:
        cfg = yaml.safe_load(f)
    cfg["SHARDS_PATH_OR_URLS"] = [f"pipe:cat {shard}"]
    cfg["K"] = [4, 4, 4, 4]                 # shorter teacher rollout; every other setting (lpips, DMD, lsgan, ...) as shipped
    cfg["BATCH_SIZE"] = 2
    cfg["NUM_STEPS"] = [1]
    cfg["CKPT_EVERY_N_STEPS"] = max(steps, 1)
    os.makedirs(os.path.join(workdir, "configs"), exist_ok=True)
    with open(os.path.join(workdir, "configs", name), "w") as f:
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
    if steps == 0 and os.environ.get("FLASH_EXAMPLE_REAL_WEIGHTS") != "1":
        # shapes only (see the module docstring); FLASH_EXAMPLE_REAL_WEIGHTS=1 materialises the ~1.9 B parameters the
        # script builds (teacher + student + CLIP + VAE + pipeline copy) — minutes on a box with slow first-touch memory
        with torch.device("meta"):
            runpy.run_path(SCRIPT, run_name="__main__")
    else:
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


FULL_SIZE = {  # script -> (EXP_NAME suffix of the log dir, keys the mapper chain must deliver, image size)
    "sdxl": ("FlashSDXL", {"image", "text", "original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple"}),
    "pixart": ("FlashPixart", {"image", "text"}),
    "sd3": ("FlashSD3", {"image", "text"}),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE))
def test_full_size_scripts_run_unchanged(name, tmp_path, monkeypatch):
    """examples/train_flash_{sdxl,pixart,sd3}.py, unmodified (see the module docstring)."""
    import torch
    ref_dir = os.environ.get("FLASH_REF_DIR", "/root/reference/examples")
    script = os.path.join(ref_dir, f"train_flash_{name}.py")
    ref_yaml = os.path.join(ref_dir, "configs", f"flash_{name}.yaml")
    if not (os.path.exists(script) and os.path.exists(ref_yaml)):
        pytest.skip("the reference tree is not on this box")
    steps = 2 if torch.cuda.is_available() else 0
    shard = str(tmp_path / "000000.tar")
    make_shard(shard)
    cfg = write_config(str(tmp_path), shard, steps, ref_yaml=ref_yaml, name=f"flash_{name}.yaml")
    assert cfg["DISTILL_LOSS_TYPE"] == "lpips" and cfg["LORA"]
    monkeypatch.chdir(tmp_path)
    for k, v in dict(SLURM_NPROCS="1", SLURM_NNODES="1", SLURM_JOB_ID="0", FLASH_MAX_STEPS=str(steps)).items():
        monkeypatch.setenv(k, v)
    compat = os.path.join(ROOT, "flash-diffusion_b200", "compat")
    monkeypatch.setattr(sys, "path", sys.path + [compat])
    if steps == 0:
        with torch.device("meta"):
            ns = runpy.run_path(script, run_name="__main__")
    else:
        ns = runpy.run_path(script, run_name="__main__")
    assert "main" in ns
    runs = sorted(os.listdir(tmp_path / "logs"))
    exp, keys = FULL_SIZE[name]
    assert len(runs) == 1 and exp in runs[0]
    summary = json.load(open(tmp_path / "logs" / runs[0] / "fit_summary.json"))
    assert summary["steps"] == steps
    if steps == 0:
        sb = summary["sanity_batch"]
        assert keys <= set(sb) and sb["image"] == [2, 3, 1024, 1024]
    else:
        assert len(summary["losses"]) == steps and all(l["loss_optimizer_0"] > 0 for l in summary["losses"])
        assert any(f.endswith("_lora.safetensors") for f in os.listdir(tmp_path / "logs" / runs[0] / "checkpoints"))
