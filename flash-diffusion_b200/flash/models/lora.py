"""LoRA injection with peft-0.9 semantics and state-dict keys (`base_layer`, `lora_A.default`, `lora_B.default`).

reference call sites: examples/train_flash_sdxl.py:210-217 (`student_unet.add_adapter(LoraConfig(r, lora_alpha=r,
init_lora_weights="gaussian", target_modules=["to_k","to_q","to_v","to_out.0"]))`),
examples/train_flash_pixart.py:237-256 (`get_peft_model`).  peft itself is not installable here; the math
`y = base(x) + (x A^T) B^T * alpha / r` is restated and executed by fd_gemm with the
LoRA product folded in as a second K segment (flash.b200.ops.LinearPack).
"""
import math

import torch.nn as nn


class LoraConfig:
    """The `peft.LoraConfig` fields the reference sets."""

    def __init__(self, r=8, lora_alpha=8, init_lora_weights=True, target_modules=None, **unused):
        self.r, self.lora_alpha, self.init_lora_weights = r, lora_alpha, init_lora_weights
        self.target_modules = list(target_modules or [])


class LoRALinear(nn.Module):
    """Parameter container for a LoRA-wrapped Linear (the forward lives in the B200 engine)."""

    def __init__(self, base: nn.Linear, r: int, lora_alpha: float, init=True):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = lora_alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        self.lora_A.to(base.weight.device)
        self.lora_B.to(base.weight.device)
        if init == "gaussian":
            nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        else:
            nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, *a, **k):
        raise RuntimeError("parameter container: executed by the B200 engine")


class LoRAConv2d(nn.Module):
    """Parameter container for a LoRA-wrapped Conv2d with peft-0.9 keys (`lora_A.default.weight` [r, Cin, kh, kw],
    `lora_B.default.weight` [Cout, r, 1, 1]).  The only convolution the example scripts' target lists reach is the DiT
    patch embedding (`pos_embed.proj` matches "proj"); its forward / backward live in
    flash.models.transformers.transformers.patch_embed."""

    def __init__(self, base: nn.Conv2d, r: int, lora_alpha: float, init=True):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = lora_alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride,
                                                         base.padding, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Conv2d(r, base.out_channels, 1, 1, bias=False)})
        self.lora_A.to(base.weight.device)
        self.lora_B.to(base.weight.device)
        if init == "gaussian":
            nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        else:
            nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, *a, **k):
        raise RuntimeError("parameter container: executed by the B200 engine")


def inject_lora(model: nn.Module, config: LoraConfig, conv_names=("pos_embed.proj",)):
    """Wrap every nn.Linear whose qualified name ends with a target suffix, and the convolutions named in `conv_names`
    when a target suffix matches them (peft wraps any matching Conv2d; the engines execute LoRA only on the DiT patch
    convolution, and no target list of the example scripts reaches another one); freeze all non-LoRA parameters
    (peft `inject_adapter_in_model` + `mark_only_lora_as_trainable`)."""
    for p in model.parameters():
        p.requires_grad = False
    n = 0
    for name, module in list(model.named_modules()):
        for child_name, child in list(module.named_children()):
            full = f"{name}.{child_name}" if name else child_name
            hit = any(full == t or full.endswith("." + t) for t in config.target_modules)
            if hit and isinstance(child, nn.Conv2d) and full not in conv_names:
                raise NotImplementedError(f"LoRA on the convolution {full!r} is not executed by the B200 engines")
            if hit and isinstance(child, (nn.Linear, nn.Conv2d)):
                cls = LoRALinear if isinstance(child, nn.Linear) else LoRAConv2d
                wrapped = cls(child, config.r, config.lora_alpha, config.init_lora_weights)
                if isinstance(module, nn.ModuleList):
                    module[int(child_name)] = wrapped
                else:
                    setattr(module, child_name, wrapped)
                n += 1
    if n == 0:
        raise ValueError(f"Target modules {config.target_modules} not found in the base model.")
    return model
