"""LoRA export / checkpointing — what makes a distilled student usable by the reference's inference
recipes (reference README.md:316-405: `PeftModel.from_pretrained(transformer, "jasperai/flash-pixart")`,
`pipe.load_lora_weights("jasperai/flash-sdxl")`, ComfyUI `models/loras/FlashSDXL.safetensors`) and the periodic
checkpoints of the training scripts (examples/train_flash_sdxl.py:438-443: `ModelCheckpoint(every_n_train_steps=...)`).

Three wire formats, all safetensors, all written from the wrapper's peft-style parameters
(`<module>.lora_A.default.weight [r, in]`, `<module>.lora_B.default.weight [out, r]`, scaling = lora_alpha / r):

  "peft"       directory: adapter_model.safetensors with keys `base_model.model.<module>.lora_A.weight` (peft drops
               the adapter name on save) + adapter_config.json — `peft.PeftModel.from_pretrained`
  "diffusers"  pytorch_lora_weights.safetensors with keys `<prefix>.<module>.lora_A.weight` (prefix `unet` or
               `transformer`) — `pipe.load_lora_weights` (peft-backend key scheme)
  "kohya"      single file with `lora_unet_<module with _>.lora_down.weight / .lora_up.weight / .alpha` — ComfyUI / A1111

peft / diffusers are not installable offline, so the key schemes are restated from their published formats; the
round trip through `load_lora` (all three formats) is tested, loading into the real libraries is not ("unpinned").
"""
import json
import os
import re
from typing import Dict

import torch

_A = re.compile(r"^(.*)\.lora_A\.default\.weight$")
_B = re.compile(r"^(.*)\.lora_B\.default\.weight$")


def lora_state_dict(denoiser: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """{'<module>.lora_A.default.weight': ..., '<module>.lora_B.default.weight': ...} (CPU, fp32, contiguous)."""
    out = {}
    for k, v in denoiser.state_dict().items():
        if _A.match(k) or _B.match(k):
            out[k] = v.detach().to("cpu", torch.float32).contiguous()
    if not out:
        raise ValueError("the module carries no LoRA adapter (add_adapter / get_peft_model first)")
    return out


def lora_meta(denoiser: torch.nn.Module):
    """(r, lora_alpha, sorted target module suffixes) read back from the injected wrappers."""
    r = alpha = None
    targets = set()
    for name, m in denoiser.named_modules():
        if hasattr(m, "base_layer") and hasattr(m, "lora_A"):
            r = m.r
            alpha = m.scaling * m.r
            targets.add(name.split(".")[-1] if not name.endswith("to_out.0") else "to_out.0")
    if r is None:
        raise ValueError("the module carries no LoRA adapter")
    return int(r), float(alpha), sorted(targets)


def _save(tensors, path, metadata=None):
    from safetensors.torch import save_file
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    save_file({k: v.contiguous() for k, v in tensors.items()}, path, metadata=metadata)


def save_lora(denoiser: torch.nn.Module, path: str, fmt: str = "diffusers", prefix: str = "unet",
              dtype: torch.dtype = torch.float16) -> str:
    """Write the adapter of `denoiser` in one of the three formats; returns the file written."""
    sd = lora_state_dict(denoiser)
    r, alpha, targets = lora_meta(denoiser)
    cast = lambda t: t.to(dtype)
    if fmt == "peft":
        os.makedirs(path, exist_ok=True)
        out = {}
        for k, v in sd.items():
            out["base_model.model." + k.replace(".default.weight", ".weight")] = cast(v)
        f = os.path.join(path, "adapter_model.safetensors")
        _save(out, f, {"format": "pt"})
        cfg = {"peft_type": "LORA", "task_type": None, "r": r, "lora_alpha": alpha, "lora_dropout": 0.0, "bias": "none",
               "target_modules": targets, "init_lora_weights": True, "inference_mode": True, "fan_in_fan_out": False,
               "base_model_name_or_path": None, "modules_to_save": None}
        with open(os.path.join(path, "adapter_config.json"), "w") as fh:
            json.dump(cfg, fh, indent=2)
        return f
    if fmt == "diffusers":
        out = {f"{prefix}." + k.replace(".default.weight", ".weight"): cast(v) for k, v in sd.items()}
        f = path if path.endswith(".safetensors") else os.path.join(path, "pytorch_lora_weights.safetensors")
        _save(out, f, {"format": "pt", "lora_alpha": str(alpha), "r": str(r)})
        return f
    if fmt == "kohya":
        out = {}
        head = "lora_unet_" if prefix == "unet" else f"lora_{prefix}_"
        for k, v in sd.items():
            ma, mb = _A.match(k), _B.match(k)
            mod = (ma or mb).group(1).replace(".", "_")
            if ma:
                out[f"{head}{mod}.lora_down.weight"] = cast(v)
                out[f"{head}{mod}.alpha"] = torch.tensor(alpha, dtype=dtype)
            else:
                out[f"{head}{mod}.lora_up.weight"] = cast(v)
        f = path if path.endswith(".safetensors") else os.path.join(path, "flash_lora_kohya.safetensors")
        _save(out, f, {"format": "pt"})
        return f
    raise ValueError(f"unknown LoRA format {fmt!r} (peft | diffusers | kohya)")


def load_lora(denoiser: torch.nn.Module, path: str, prefix: str = "unet") -> int:
    """Load an adapter written by `save_lora` (any format) into an adapter-carrying denoiser; returns the number of
    tensors loaded.  Keys are mapped back to `<module>.lora_{A,B}.default.weight`."""
    from safetensors.torch import load_file
    if os.path.isdir(path):
        for cand in ("adapter_model.safetensors", "pytorch_lora_weights.safetensors", "flash_lora_kohya.safetensors"):
            if os.path.exists(os.path.join(path, cand)):
                path = os.path.join(path, cand)
                break
    raw = load_file(path)
    own = denoiser.state_dict()
    under = {k.replace(".", "_"): k for k in {(_A.match(k) or _B.match(k)).group(1) for k in own if _A.match(k) or _B.match(k)}}
    mapped = {}
    for k, v in raw.items():
        if k.endswith(".alpha"):
            continue
        if k.startswith("base_model.model."):
            k2 = k[len("base_model.model."):].replace(".lora_A.weight", ".lora_A.default.weight").replace(
                ".lora_B.weight", ".lora_B.default.weight")
        elif k.startswith("lora_"):
            body, kind = k.rsplit(".lora_", 1)
            body = body.split("_", 2)[2] if body.startswith("lora_unet_") else body.split("_", 2)[2]
            if body not in under:
                raise KeyError(f"kohya key {k!r} matches no adapted module")
            k2 = under[body] + (".lora_A.default.weight" if kind.startswith("down") else ".lora_B.default.weight")
        else:
            k2 = k.split(".", 1)[1] if k.split(".", 1)[0] in (prefix, "unet", "transformer") else k
            k2 = k2.replace(".lora_A.weight", ".lora_A.default.weight").replace(".lora_B.weight", ".lora_B.default.weight")
        if k2 not in own:
            raise KeyError(f"adapter key {k!r} -> {k2!r} not found in the module")
        mapped[k2] = v.to(own[k2].dtype)
    missing = [k for k in own if (_A.match(k) or _B.match(k)) and k not in mapped]
    if missing:
        raise KeyError(f"adapter file lacks {len(missing)} tensors, e.g. {missing[0]}")
    with torch.no_grad():
        params = dict(denoiser.named_parameters())
        for k, v in mapped.items():
            params[k].copy_(v.to(params[k].device))
    return len(mapped)


def merge_lora_into_base(denoiser: torch.nn.Module) -> int:
    """W <- W + scaling * B A for every adapted Linear and B <- 0 (the inference-time "fused" LoRA of
    `pipe.fuse_lora()`); returns the number of merged layers."""
    n = 0
    with torch.no_grad():
        for m in denoiser.modules():
            if hasattr(m, "base_layer") and hasattr(m, "lora_A") and isinstance(m.base_layer, torch.nn.Linear):
                A, B = m.lora_A["default"].weight, m.lora_B["default"].weight
                m.base_layer.weight.add_(m.scaling * (B.float() @ A.float()).to(m.base_layer.weight.dtype))
                B.zero_()
                n += 1
    return n


class ModelCheckpoint:
    """Every `every_n_train_steps` training steps: `<dirpath>/<filename>.ckpt` (trainable parameters + optimizer states
    + step; `save_frozen=True` adds the frozen teacher/base weights like Lightning's full `state_dict`) and, next to it,
    the student's LoRA in diffusers format.  Mirrors the `pytorch_lightning.callbacks.ModelCheckpoint` arguments the
    example scripts pass (examples/train_flash_sdxl.py:438-443)."""

    def __init__(self, dirpath: str, filename: str = "{step}", every_n_train_steps: int = 1000, save_top_k: int = -1,
                 save_frozen: bool = False, lora_format: str = "diffusers", **unused):
        self.dirpath, self.filename, self.every = dirpath, filename, max(1, int(every_n_train_steps))
        self.save_frozen, self.lora_format = save_frozen, lora_format
        self.saved = []

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        step = getattr(trainer, "global_step", batch_idx + 1)
        if step % self.every == 0:
            self.save(pl_module, step)

    def save(self, pipeline, step: int) -> str:
        os.makedirs(self.dirpath, exist_ok=True)
        name = self.filename.format(step=f"step={step}") if "{step}" in self.filename else self.filename
        path = os.path.join(self.dirpath, name + ".ckpt")
        model = pipeline.model
        keep = {n for n, p in model.named_parameters() if p.requires_grad}
        sd = {("model." + k): v.detach().cpu() for k, v in model.state_dict().items()
              if self.save_frozen or k in keep or "lora_" in k or k.startswith("discriminator")}
        opts = [o.state_dict() for o in (pipeline.optims or [])]
        torch.save({"state_dict": sd, "optimizer_states": opts, "global_step": step}, path)
        student = getattr(model, "student_denoiser", None)
        if student is not None:
            try:
                save_lora(student, os.path.join(self.dirpath, name + "_lora.safetensors"), fmt=self.lora_format)
            except ValueError:
                pass        # full fine-tuning (no adapter): the .ckpt holds the student
        self.saved.append(path)
        return path
