"""LoRA export / checkpoint round trips (reference README.md:316-405, train_flash_sdxl.py:438-443)."""
import os

import pytest
import torch

from flash.models.lora import LoraConfig
from flash.models.unets import DiffusersUNet2DCondWrapper
from flash.recipes import TINY_UNET_KWARGS
from flash.trainer.export import ModelCheckpoint, load_lora, lora_meta, lora_state_dict, merge_lora_into_base, save_lora


def _student(seed=0):
    torch.manual_seed(seed)
    net = DiffusersUNet2DCondWrapper(**TINY_UNET_KWARGS)
    net.add_adapter(LoraConfig(r=8, lora_alpha=16, init_lora_weights="gaussian",
                               target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "lora_B" in n:
                p.normal_(0, 0.05)
    return net


@pytest.mark.parametrize("fmt", ["peft", "diffusers", "kohya"])
def test_save_load_round_trip(tmp_path, fmt):
    a, b = _student(0), _student(1)
    f = save_lora(a, str(tmp_path / "out"), fmt=fmt, dtype=torch.float32)
    assert os.path.exists(f)
    from safetensors.torch import load_file
    keys = list(load_file(f))
    if fmt == "peft":
        assert all(k.startswith("base_model.model.") and ".default." not in k for k in keys)
        assert os.path.exists(tmp_path / "out" / "adapter_config.json")
    elif fmt == "diffusers":
        assert all(k.startswith("unet.") and k.endswith((".lora_A.weight", ".lora_B.weight")) for k in keys)
    else:
        assert any(k.endswith(".lora_down.weight") for k in keys) and any(k.endswith(".alpha") for k in keys)
        assert all(k.startswith("lora_unet_") for k in keys)
    n = load_lora(b, str(tmp_path / "out"))
    sa, sb = lora_state_dict(a), lora_state_dict(b)
    assert n == len(sa) and all(torch.equal(sa[k], sb[k]) for k in sa)


def test_meta_and_merge():
    net = _student()
    r, alpha, targets = lora_meta(net)
    assert (r, alpha) == (8, 16.0) and targets == ["to_k", "to_out.0", "to_q", "to_v"]
    lin = next(m for m in net.modules() if hasattr(m, "base_layer"))
    x = torch.randn(3, lin.base_layer.in_features)
    A, B = lin.lora_A["default"].weight, lin.lora_B["default"].weight
    want = x @ lin.base_layer.weight.t() + (x @ A.t()) @ B.t() * lin.scaling
    assert merge_lora_into_base(net) > 0
    got = x @ lin.base_layer.weight.t()
    assert torch.allclose(got, want, atol=1e-5) and float(B.abs().max()) == 0


def test_model_checkpoint_callback(tmp_path):
    class Pipe:
        pass
    pipe = Pipe()
    pipe.model = torch.nn.Module()
    pipe.model.student_denoiser = _student()
    pipe.model.discriminator = torch.nn.Linear(4, 1)
    pipe.optims = [torch.optim.AdamW([p for p in pipe.model.parameters() if p.requires_grad], lr=1e-3)]
    cb = ModelCheckpoint(dirpath=str(tmp_path), filename="{step}", every_n_train_steps=2, save_top_k=-1)

    class T:
        global_step = 0
    t = T()
    for i in range(4):
        t.global_step = i + 1
        cb.on_train_batch_end(t, pipe, None, None, i)
    assert [os.path.basename(p) for p in cb.saved] == ["step=2.ckpt", "step=4.ckpt"]
    ck = torch.load(cb.saved[-1])
    assert ck["global_step"] == 4 and any("lora_A" in k for k in ck["state_dict"]) and len(ck["optimizer_states"]) == 1
    assert not any(k.endswith("conv_in.weight") for k in ck["state_dict"])          # frozen base not duplicated
    assert os.path.exists(tmp_path / "step=4_lora.safetensors")
